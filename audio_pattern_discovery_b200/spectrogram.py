"""The NDSequence layout contract of the reference (src/spectrogram.rs:13-24): the
container the DTW path consumes.  Only the members the hot path touches are
mirrored -- n_bins, frames, vec(t) (99-101), len() (152-154); cepstrum extraction,
VAT slicing and image export are outside SURVEY.md section 8.
"""
import numpy as np


class NDSequence:
    def __init__(self, n_bins, frames, audio_id=0):
        self.n_bins = int(n_bins)
        self.frames = np.ascontiguousarray(frames, dtype=np.float32).ravel()
        self.audio_id = int(audio_id)
        if self.n_bins <= 0:
            raise ValueError("n_bins must be positive")

    @staticmethod
    def from_array(a, audio_id=0):
        a = np.ascontiguousarray(a, dtype=np.float32)
        if a.ndim == 1:
            a = a.reshape(-1, 1)
        return NDSequence(a.shape[1], a, audio_id)

    def vec(self, t):
        """src/spectrogram.rs:99-101"""
        return self.frames[t * self.n_bins:(t + 1) * self.n_bins]

    def len(self):
        """src/spectrogram.rs:152-154 (integer division: a ragged tail is ignored)"""
        return self.frames.size // self.n_bins

    def __len__(self):
        return self.len()

    def as_array(self):
        t = self.len()
        return self.frames[:t * self.n_bins].reshape(t, self.n_bins)
