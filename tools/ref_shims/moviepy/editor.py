class AudioFileClip:   # multimodal_datasets.py:14 (data loading: not on the denoising path)
    def __init__(self, *a, **k):
        raise RuntimeError("moviepy stand-in: decoding audio files is not supported")
