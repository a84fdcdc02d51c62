set -x
APD_WIDE=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_c1_pipeline.py -m gpu -q -x > gpurun_out/r2t_pytest_wide.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2t_pytest_wide.log
tail -3 gpurun_out/r2t_pytest_wide.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "width or golden or ring" > gpurun_out/r2t_pytest_widths.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2t_pytest_widths.log
tail -3 gpurun_out/r2t_pytest_widths.log
B="--seqs 4000 --steps 3 --warmup 2 --no-cpu --e2e-steps 1 --other-mode-steps 1"
timeout 200 python bench.py $B > gpurun_out/r2t_c3_4000_default.json 2> gpurun_out/r2t_d.err; echo "rc=$?"
APD_WIDE=1 timeout 200 python bench.py $B > gpurun_out/r2t_c3_4000_wide2row.json 2> gpurun_out/r2t_w.err; echo "rc=$?"
APD_WIDE=1 timeout 200 python bench.py --workload C2 --steps 3 --warmup 2 --no-cpu --e2e-steps 1 --other-mode-steps 1 > gpurun_out/r2t_c2_wide2row.json 2> gpurun_out/r2t_c2w.err; echo "rc=$?"
APD_WIDE=1 timeout 300 python bench.py --workload C4 --seqs 5000 --steps 2 --warmup 1 --no-cpu --e2e-steps 1 --other-mode-steps 1 > gpurun_out/r2t_c4_5000_wide2row.json 2> gpurun_out/r2t_c4w.err; echo "rc=$?"
