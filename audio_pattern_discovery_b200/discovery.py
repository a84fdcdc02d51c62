"""Mirror of the reference's parameter struct (src/discovery.rs:7-46).  All fields of the
TOML are kept so a reference Discovery.toml loads unchanged; only the six DTW /
clustering fields are used by this package.
"""
import numpy as np

from .alignments import AlignmentParams

_FIELDS = (("dft_win", int), ("dft_step", int), ("ceps_filter", int), ("vat_moving", int),
           ("vat_percentile", float), ("vat_min_len", int), ("alignment_workers", int),
           ("clustering_percentile", float), ("warping_band_percentage", float),
           ("insertion_penalty", float), ("deletion_penalty", float), ("match_penalty", float),
           ("auto_encoder", int), ("learning_rate", float), ("epochs", int), ("epoch_drop", float),
           ("drop", float))

# project/config/Discovery.toml:1-22
_DEFAULTS = dict(dft_win=256, dft_step=128, ceps_filter=32, auto_encoder=10, learning_rate=0.1,
                 epochs=25, epoch_drop=5.0, drop=0.5, vat_moving=15, vat_percentile=0.95,
                 vat_min_len=150, warping_band_percentage=1.0, insertion_penalty=1.0,
                 deletion_penalty=1.0, match_penalty=1.0, alignment_workers=4,
                 clustering_percentile=0.05)


def f32_as_usize(v):
    """Rust `f32 as usize`: truncation toward zero, saturating, NaN -> 0."""
    v = float(v)
    if v != v or v <= 0.0:
        return 0
    if v >= 18446744073709551616.0:
        return (1 << 64) - 1
    return int(v)


class Discovery:
    def __init__(self, **kw):
        vals = dict(_DEFAULTS)
        unknown = set(kw) - set(vals)
        if unknown:
            raise TypeError("unknown Discovery field(s): %s" % sorted(unknown))
        vals.update(kw)
        for name, typ in _FIELDS:
            setattr(self, name, typ(vals[name]))

    @staticmethod
    def from_toml(file):
        """src/discovery.rs:29-36: every field is required, like serde's Deserialize."""
        import tomllib
        with open(file, "rb") as f:
            conf = tomllib.load(f)
        missing = [n for n, _ in _FIELDS if n not in conf]
        if missing:
            raise KeyError("missing field `%s`" % missing[0])
        return Discovery(**{n: conf[n] for n, _ in _FIELDS})

    def alignment_params(self, n_size):
        """src/discovery.rs:38-45: the band is an f32 product truncated to usize."""
        band = f32_as_usize(np.float32(self.warping_band_percentage) * np.float32(n_size))
        return AlignmentParams(band, self.insertion_penalty, self.deletion_penalty, self.match_penalty)

    def clone(self):
        return Discovery(**{n: getattr(self, n) for n, _ in _FIELDS})
