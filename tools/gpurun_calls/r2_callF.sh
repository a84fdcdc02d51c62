set -x
APD_WIDE=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_c1_pipeline.py -m gpu -q -x > gpurun_out/r2f_pytest_wide.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest_wide.log
tail -4 gpurun_out/r2f_pytest_wide.log
W=$PWD/audio_pattern_discovery_b200/libapd_b200.wedge.so
B="--steps 3 --warmup 2 --no-cpu --no-parity --e2e-steps 1"
timeout 200 python bench.py --seqs 4000 $B > gpurun_out/r2f_c3_4000_default.json 2> gpurun_out/r2f_c3_4000_default.err; echo "rc=$?"
APD_WIDE=1 timeout 200 python bench.py --seqs 4000 $B > gpurun_out/r2f_c3_4000_wide.json 2> gpurun_out/r2f_c3_4000_wide.err; echo "rc=$?"
timeout 200 python bench.py --workload C2 $B > gpurun_out/r2f_c2_default.json 2> gpurun_out/r2f_c2_default.err; echo "rc=$?"
APD_LIB_PATH=$W timeout 200 python bench.py --workload C2 $B > gpurun_out/r2f_c2_wedge.json 2> gpurun_out/r2f_c2_wedge.err; echo "rc=$?"
APD_WIDE=1 timeout 200 python bench.py --workload C2 $B > gpurun_out/r2f_c2_wide.json 2> gpurun_out/r2f_c2_wide.err; echo "rc=$?"
timeout 300 python bench.py --workload C4 --seqs 5000 $B > gpurun_out/r2f_c4_5000_default.json 2> gpurun_out/r2f_c4_5000_default.err; echo "rc=$?"
APD_WIDE=1 timeout 300 python bench.py --workload C4 --seqs 5000 $B > gpurun_out/r2f_c4_5000_wide.json 2> gpurun_out/r2f_c4_5000_wide.err; echo "rc=$?"
timeout 100 python bench.py --workload C1ref $B > gpurun_out/r2f_c1ref_default.json 2> gpurun_out/r2f_c1ref_default.err; echo "rc=$?"
APD_WIDE=1 python tools/profile_case.py --workload C3 --n 1000 --mode strict > gpurun_out/r2f_prof_c3_wide_plain.log 2>&1 && \
APD_WIDE=1 ncu --set full --clock-control none --import-source on -k regex:dtw_units -c 1 -o gpurun_out/r2f_ncu_c3_strict_wide -f python tools/profile_case.py --workload C3 --n 1000 --mode strict > gpurun_out/r2f_ncu_c3_wide.log 2>&1; echo "ncu rc=$?"
