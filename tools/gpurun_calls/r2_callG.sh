set -x
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2g_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_pytest_gpu.log
tail -4 gpurun_out/r2g_pytest_gpu.log
B="--no-cpu --e2e-steps 2"
timeout 200 python bench.py --workload C2 --steps 5 --warmup 3 $B > gpurun_out/r2g_bench_C2_n1.json 2> gpurun_out/r2g_bench_C2_n1.err; echo "rc=$?"
timeout 100 python bench.py --workload C1ref --steps 5 --warmup 3 $B > gpurun_out/r2g_bench_C1ref_n1.json 2> gpurun_out/r2g_bench_C1ref_n1.err; echo "rc=$?"
timeout 400 python bench.py --steps 5 --warmup 3 > gpurun_out/r2g_bench_C3_n1.json 2> gpurun_out/r2g_bench_C3_n1.err; echo "rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2g_bench_C3_reference.json 2> gpurun_out/r2g_bench_C3_reference.err; echo "rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2g_launches_bench_c3.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-parity --other-mode-steps 0 --e2e-steps 1 > gpurun_out/r2g_ncu_launches.log 2>&1; echo "ncu rc=$?"
python tools/profile_case.py --workload C3 --n 1000 --mode fast > gpurun_out/r2g_prof_c3_fast_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:dtw_units -c 1 -o gpurun_out/r2g_ncu_c3_fast -f python tools/profile_case.py --workload C3 --n 1000 --mode fast > gpurun_out/r2g_ncu_c3_fast.log 2>&1; echo "ncu rc=$?"
python tools/profile_case.py --workload C2 --n 2000 --mode strict > gpurun_out/r2g_prof_c2_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:dtw_units -c 3 -o gpurun_out/r2g_ncu_c2_strict -f python tools/profile_case.py --workload C2 --n 2000 --mode strict > gpurun_out/r2g_ncu_c2.log 2>&1; echo "ncu rc=$?"
