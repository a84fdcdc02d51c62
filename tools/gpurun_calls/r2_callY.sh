set -x
timeout 900 python bench.py --workload C4 --steps 1 --warmup 1 --e2e-steps 1 --other-mode-steps 1 --no-cpu > gpurun_out/r2y_bench_C4_n1.json 2> gpurun_out/r2y_c4.err; echo "rc=$?"
timeout 600 python bench.py --workload C5 --steps 1 --warmup 1 --e2e-steps 1 --other-mode-steps 1 --no-cpu > gpurun_out/r2y_bench_C5_n1.json 2> gpurun_out/r2y_c5.err; echo "rc=$?"
timeout 100 python bench.py --workload C1ref --steps 5 --warmup 3 --e2e-steps 2 --no-cpu > gpurun_out/r2y_bench_C1ref_n1.json 2> gpurun_out/r2y_c1.err; echo "rc=$?"
