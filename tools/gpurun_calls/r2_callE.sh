set -x
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2e_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest_gpu.log
tail -4 gpurun_out/r2e_pytest_gpu.log
C=$PWD/audio_pattern_discovery_b200/libapd_b200.compact.so
for w in C2 C1ref; do
  timeout 200 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu > gpurun_out/r2e_${w}_default.json 2> gpurun_out/r2e_${w}_default.err; echo "rc=$?"
  APD_LIB_PATH=$C timeout 200 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu > gpurun_out/r2e_${w}_compact.json 2> gpurun_out/r2e_${w}_compact.err; echo "rc=$?"
done
APD_LIB_PATH=$C timeout 200 python bench.py --seqs 4000 --steps 3 --warmup 2 --no-cpu --no-parity --e2e-steps 1 > gpurun_out/r2e_c3_4000_compact.json 2> gpurun_out/r2e_c3_4000_compact.err; echo "rc=$?"
python tools/tensor_core_score_error.py > gpurun_out/r2_tensor_core_score_error.json 2> gpurun_out/r2_tensor_core_score_error.err; echo "rc=$?"
timeout 600 python bench.py --workload C5 --steps 1 --warmup 1 --e2e-steps 1 --other-mode-steps 1 --no-cpu > gpurun_out/r2e_c5_full.json 2> gpurun_out/r2e_c5_full.err; echo "rc=$?"
