"""blobfile stand-in for local paths (reference call sites: dist_util.py:59-66, multimodal_train_util.py:212-537,
multimodal_datasets.py:107-113)."""
import os

join = os.path.join
dirname = os.path.dirname
basename = os.path.basename
exists = os.path.exists
isdir = os.path.isdir
listdir = os.listdir


def BlobFile(path, mode="r", **_):
    return open(path, mode)


def makedirs(path):
    os.makedirs(path, exist_ok=True)
