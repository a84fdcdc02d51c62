set -x
timeout 400 python bench.py --steps 3 --warmup 2 --no-cpu --e2e-steps 1 > gpurun_out/r2i_bench_C3_n1.json 2> gpurun_out/r2i_bench_C3_n1.err; echo "rc=$?"
T=$PWD/audio_pattern_discovery_b200/libapd_b200.tma.so
APD_LIB_PATH=$T timeout 200 python bench.py --seqs 4000 --steps 3 --warmup 2 --no-cpu --e2e-steps 1 > gpurun_out/r2i_bench_C3_4000_tma_variant.json 2> gpurun_out/r2i_tma.err; echo "rc=$?"
timeout 200 python bench.py --seqs 4000 --steps 3 --warmup 2 --no-cpu --e2e-steps 1 > gpurun_out/r2i_bench_C3_4000_default.json 2> gpurun_out/r2i_def.err; echo "rc=$?"
APD_LIB_PATH=$T python tools/profile_case.py --workload C3 --n 1000 --mode strict > gpurun_out/r2i_prof_tma_plain.log 2>&1 && \
APD_LIB_PATH=$T ncu --set full --clock-control none --import-source on -k regex:dtw_units -c 1 -o gpurun_out/r2i_ncu_c3_strict_tma_variant -f python tools/profile_case.py --workload C3 --n 1000 --mode strict > gpurun_out/r2i_ncu_tma.log 2>&1; echo "ncu rc=$?"
python tools/profile_case.py --workload C3 --n 1000 --mode strict > gpurun_out/r2i_prof_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:dtw_units -c 1 -o gpurun_out/r2i_ncu_c3_strict -f python tools/profile_case.py --workload C3 --n 1000 --mode strict > gpurun_out/r2i_ncu.log 2>&1; echo "ncu rc=$?"
timeout 300 python bench.py --workload C4 --seqs 5000 --steps 2 --warmup 1 --no-cpu --e2e-steps 1 > gpurun_out/r2i_bench_C4_5000.json 2> gpurun_out/r2i_c4.err; echo "rc=$?"
