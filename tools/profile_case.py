"""One all-pairs call of a reduced workload through the public interface -- the command
profiled with ncu (see profiles/README.md).  Not a benchmark: prints library-side stats only.

    python tools/profile_case.py [--workload C3] [--n 640] [--mode fast|strict] [--reps 1]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from audio_pattern_discovery_b200 import APD_MODE_FAST, APD_MODE_STRICT, Context, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="C3")
    ap.add_argument("--n", type=int, default=640)
    ap.add_argument("--mode", default="strict")
    ap.add_argument("--reps", type=int, default=1)
    a = ap.parse_args()
    c, seqs, _ = synth.make_config(a.workload, a.n)
    mode = APD_MODE_STRICT if a.mode == "strict" else APD_MODE_FAST
    with Context(0) as ctx:
        ctx.set_sequences(seqs)
        for _ in range(a.reps):
            ctx.align_all(c["pct"], *c["weights"], mode=mode)
            st = ctx.stats()
            print("%s n=%d %s: kernel %.3f ms, %.1f GCUPS (reference cells %d, computed %d), %d launches"
                  % (a.workload, a.n, a.mode, st["kernel_ms"], st["cells_reference"] / st["kernel_ms"] / 1e6,
                     st["cells_reference"], st["cells_computed"], st["kernel_launches"]))


if __name__ == "__main__":
    main()
