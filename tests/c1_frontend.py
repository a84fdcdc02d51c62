"""Input generator for BASELINE.json config 1 -- TEST INFRASTRUCTURE ONLY, OUT OF SCOPE FOR PARITY.

A numpy restatement of the reference's front-end, used only to produce
DTW inputs with the value distribution the real pipeline feeds the hot path: synthetic
chirp / whistle recordings -> cepstrum (src/spectrogram.rs:31-94: `0.54 + 0.46 cos` window, 256-point
DFT magnitude of the first 128 bins, 8-tap triangle filter stride 4 -> 30 values, ln(x + 1e-6),
DCT-I, drop 4 coefficients, mean-centre -> 26 bins) -> "interesting" slices by smoothed per-frame
deviation (src/spectrogram.rs:174-216) -> one-hidden-layer auto-encoder, per-frame SGD
(src/neural.rs:74-94) -> embeddings `sigmoid(xW + b) * 255`, per-frame z-score with sigma floored
at 1 (src/neural.rs:55-71) -> sequences of dim 10.

Deviations, all irrelevant to the DTW boundary where parity is asserted (same frames in -> same
matrix / merges out): numpy's FFT and an explicit DCT-I matrix replace rustfft / rustdct (unpinned
upstream versions), the auto-encoder is seeded (upstream uses thread_rng) and trained for
`epochs` passes over a frame subsample, and `vat_percentile` defaults to 0.55 instead of 0.95
because the synthetic recordings are seconds, not minutes, long.
"""
import numpy as np

F = np.float32


def synth_recording(rng, seconds=24.0, rate=22050, n_events=11):
    n = int(seconds * rate)
    t = np.arange(n) / rate
    x = rng.normal(0.0, 30.0, size=n)
    starts = np.sort(rng.uniform(0.3, seconds - 2.2, size=n_events))
    kinds = rng.integers(0, 4, size=n_events)
    for s, k in zip(starts, kinds):
        dur = rng.uniform(0.9, 1.7)
        m = (t >= s) & (t < s + dur)
        tt = t[m] - s
        f0 = (2000, 9000, 3000, 6000)[k] * rng.uniform(0.9, 1.1)
        f1 = (9000, 2500, 3000, 11000)[k] * rng.uniform(0.9, 1.1)
        if k == 2:   # whistle with vibrato
            phase = 2 * np.pi * (f0 * tt + 120.0 * np.sin(2 * np.pi * 6.0 * tt) / (2 * np.pi * 6.0))
        elif k == 3:  # quadratic FM
            phase = 2 * np.pi * (f0 * tt + (f1 - f0) * tt ** 3 / (3 * dur ** 2))
        else:        # linear chirp up / down
            phase = 2 * np.pi * (f0 * tt + 0.5 * (f1 - f0) * tt ** 2 / dur)
        env = np.hanning(m.sum()) ** 0.5
        x[m] += 6000.0 * env * np.sin(phase)
    return np.clip(x, -32768, 32767).astype(np.int16)


def cepstrum(raw, fft_size=256, fft_step=128, filter_size=32):
    """src/spectrogram.rs:31-94 -> (frames, 26) float32."""
    x = raw.astype(F)
    win = (F(0.54) + F(0.46) * np.cos(F(2 * np.pi) * np.arange(fft_size, dtype=F) / F(fft_size))).astype(F)
    ln = fft_size // filter_size                      # 8
    tri = np.zeros(ln, dtype=F)
    center = (ln - 1) // 2
    for i in range(center + 1):
        tri[i] = tri[ln - 1 - i] = F(i) / F(ln)
    ends = np.arange(fft_size, len(x), fft_step)
    frames = np.stack([x[e - fft_size:e] * win for e in ends]).astype(F)
    mag = np.abs(np.fft.fft(frames, axis=1)[:, :fft_size // 2]).astype(F)
    idx = np.arange(ln, mag.shape[1], ln // 2)       # convolve(result, triag, step = len / 2)
    conv = np.stack([mag[:, i - ln:i] @ tri for i in idx], axis=1).astype(F)
    logc = np.log(conv + F(1e-6)).astype(F)
    n = logc.shape[1]                                 # 30
    k = np.arange(n)
    dct1 = np.cos(np.pi * np.outer(k, k) / (n - 1))   # unnormalised DCT-I
    dct1[:, 0] *= 0.5
    dct1[:, -1] *= 0.5
    ceps = (logc.astype(np.float64) @ dct1.T).astype(F)[:, 4:]
    return (ceps - ceps.mean(axis=1, keepdims=True)).astype(F)


def interesting_ranges(ceps, moving=15, perc=0.55, min_len=150):
    """src/spectrogram.rs:174-216."""
    dev = ceps.std(axis=1)
    mov = np.zeros(len(dev), dtype=F)
    for i in range(moving, len(dev)):
        mov[i] = dev[i - moving:i].mean()
    th = np.sort(mov)[int(F(len(mov)) * F(perc))]
    out, start, recording = [], 0, True
    for i, v in enumerate(mov):
        if v >= th and not recording:
            start, recording = i, True
        if v < th and recording:
            recording = False
            if i - start > min_len:
                out.append((start, i))
    return out


class AutoEncoder:
    """src/neural.rs:13-95 with a seeded initialisation."""

    def __init__(self, rng, d_in=26, latent=10):
        self.we = rng.normal(0, 0.1, size=(d_in, latent)).astype(F)
        self.wd = rng.normal(0, 0.1, size=(latent, d_in)).astype(F)
        self.be = rng.normal(0, 0.1, size=(1, latent)).astype(F)
        self.bd = rng.normal(0, 0.1, size=(1, d_in)).astype(F)

    @staticmethod
    def _sig(z):
        return (1.0 / (1.0 + np.exp(-np.clip(z, -80.0, 80.0)))).astype(F)

    def take_step(self, x, alpha):
        x = x.reshape(1, -1)
        latent = x @ self.we + self.be                # the decoder consumes the PRE-activation latent
        la = self._sig(latent)
        act = self._sig(latent @ self.wd + self.bd)
        d_out = -(x - act) * act * (1 - act)
        d_dec = (d_out @ self.wd.T) * la * (1 - la)
        self.wd -= alpha * (latent.T @ d_out)
        self.we -= alpha * (x.T @ d_dec)
        self.be -= alpha * d_dec
        self.bd -= alpha * d_out

    def predict(self, frames):
        p = self._sig(frames @ self.we + self.be) * F(255.0)
        mu = p.mean(axis=1, keepdims=True)
        sd = np.maximum(p.std(axis=1, keepdims=True), F(1.0))
        return ((p - mu) / sd).astype(F)


def make_c1_sequences(n_files=32, seed=1001, epochs=2, train_frames=6000, vat_percentile=0.55):
    """-> list of (T, 10) float32 embedding sequences (the `signals` of src/main.rs:150-161)."""
    rng = np.random.default_rng(seed)
    slices = []
    for _ in range(n_files):
        c = cepstrum(synth_recording(rng))
        for a, b in interesting_ranges(c, perc=vat_percentile):
            slices.append(c[a:b])
    ae = AutoEncoder(rng)
    allf = np.concatenate(slices)
    alpha = F(0.1)
    for _ in range(epochs):
        for i in rng.permutation(len(allf))[:train_frames]:
            ae.take_step(allf[i], alpha)
    return [ae.predict(s) for s in slices]
