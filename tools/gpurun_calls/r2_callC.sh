set -x
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2c_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest_gpu.log
tail -15 gpurun_out/r2c_pytest_gpu.log
timeout 400 python bench.py --workload C5 --seqs 300 --steps 1 --warmup 1 --e2e-steps 1 --other-mode-steps 0 --no-cpu > gpurun_out/r2c_c5_300.json 2> gpurun_out/r2c_c5_300.err; echo "rc=$?"
for i in 1 2 3; do timeout 200 python bench.py --workload C2 --steps 3 --warmup 3 --no-cpu --no-parity --other-mode-steps 0 > gpurun_out/r2c_c2_$i.json 2> gpurun_out/r2c_c2_$i.err; done
tail -n 3 gpurun_out/r2c_c5_300.err
