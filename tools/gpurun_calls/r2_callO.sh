set -x
timeout 600 python -m pytest tests/test_gpu_paths.py tests/test_gpu_parity.py -m gpu -q -x -k "path or pair or golden or c5" > gpurun_out/r2o_pytest_paths.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2o_pytest_paths.log
tail -3 gpurun_out/r2o_pytest_paths.log
cat > /tmp/k2_case.py <<'PY'
import sys, numpy as np
sys.path.insert(0, '.')
from audio_pattern_discovery_b200 import Context, synth, APD_MODE_FAST
c, seqs, _ = synth.make_config("C5", 64)
pairs = [(i, (i * 7 + 3) % 64) for i in range(64) if i != (i * 7 + 3) % 64] * 20
with Context(0) as ctx:
    ctx.set_sequences(seqs)
    for mode in (0, 0, APD_MODE_FAST):
        s, p, l = ctx.align_pairs(pairs, c["pct"], mode=mode, want_paths=True, path_cap=8200)
        print("K2 mode %d: %d pairs of 4096^2, kernels %.1f ms" % (mode, len(pairs), ctx.stats()["path_ms"]))
PY
python /tmp/k2_case.py
