import numpy as np


class ImageSequenceClip:
    def __init__(self, sequence, fps=None, **_):
        self.frames = np.stack([np.asarray(f) for f in sequence])
        self.fps = fps
        self.audio = None

    def set_audio(self, audio_clip):
        self.audio = audio_clip
        return self

    def write_videofile(self, path, fps=None, audio=True, audio_fps=None, **_):
        extra = {} if self.audio is None else {"audio": self.audio.array, "audio_fps": audio_fps or self.audio.fps}
        np.savez_compressed(path + ".npz", frames=self.frames, fps=fps or self.fps, **extra)
        open(path, "wb").close()
