"""moviepy stand-in (reference call sites: common.py:28-53, multimodal_train_util.py:15-16): clips keep their arrays and
'write' them as compressed .npz next to the requested path — encoding mp4 / wav is IO outside the denoising path."""
