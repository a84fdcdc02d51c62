set -x
cat > /tmp/k2_case.py <<'PY'
import sys, numpy as np
sys.path.insert(0, '.')
from audio_pattern_discovery_b200 import Context, synth
c, seqs, _ = synth.make_config("C5", 64)
pairs = [(i, (i * 7 + 3) % 64) for i in range(64) if i != (i * 7 + 3) % 64] * 20
with Context(0) as ctx:
    ctx.set_sequences(seqs)
    for _ in range(2):
        s, p, l = ctx.align_pairs(pairs, c["pct"], want_paths=True, path_cap=8200)
        print("K2: %d pairs of 4096^2, kernels %.1f ms" % (len(pairs), ctx.stats()["path_ms"]))
PY
python /tmp/k2_case.py > gpurun_out/r2n_k2_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:pair_ -c 2 -o gpurun_out/r2n_ncu_k2_strict -f python /tmp/k2_case.py > gpurun_out/r2n_ncu_k2.log 2>&1; echo "ncu rc=$?"
cat gpurun_out/r2n_k2_plain.log
