set -x
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2u_pytest_gpu_1gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2u_pytest_gpu_1gpu.log
tail -4 gpurun_out/r2u_pytest_gpu_1gpu.log
python -c "import __graft_entry__ as g; g.smoke()"; echo "smoke rc=$?"
timeout 200 python bench.py --workload C2 --steps 5 --warmup 3 --no-cpu --e2e-steps 2 > gpurun_out/r2u_bench_C2_n1.json 2> gpurun_out/r2u_c2.err; echo "rc=$?"
python tools/profile_case.py --workload C2 --n 2000 --mode strict > gpurun_out/r2u_prof_c2_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:dtw_units -c 3 -o gpurun_out/r2u_ncu_c2_strict -f python tools/profile_case.py --workload C2 --n 2000 --mode strict > gpurun_out/r2u_ncu_c2.log 2>&1; echo "ncu rc=$?"
