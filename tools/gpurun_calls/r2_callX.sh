set -x
python -c "import __graft_entry__ as g; g.smoke()"; echo "smoke rc=$?"
python - <<'PY'
import time, sys
sys.path.insert(0, '.')
import numpy as np
from audio_pattern_discovery_b200 import Context
seqs = [np.random.default_rng(k).normal(size=(100, 20)).astype(np.float32) for k in range(64)]
ts = []
for k in range(8):
    t0 = time.perf_counter(); c = Context(0); t1 = time.perf_counter(); c.set_sequences(seqs); m = c.align_all(0.1); t2 = time.perf_counter(); c.close()
    ts.append((round((t1 - t0) * 1e3, 2), round((t2 - t1) * 1e3, 2)))
print("create ms / first set+align ms:", ts)
PY
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_group.py -m gpu -q -x > gpurun_out/r2x_pytest.log 2>&1; tail -2 gpurun_out/r2x_pytest.log
