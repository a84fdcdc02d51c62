set -x
python tools/tensor_core_score_error.py > gpurun_out/r2_tensor_core_score_error.json 2> gpurun_out/r2_tensor_core_score_error.err; echo "rc=$?"
tail -8 gpurun_out/r2_tensor_core_score_error.err
APD_DEBUG=1 timeout 300 python bench.py --workload C2 --steps 2 --warmup 2 --no-cpu --no-parity --other-mode-steps 0 > gpurun_out/r2d_c2_dbg.json 2> gpurun_out/r2d_c2_dbg.err; echo "rc=$?"
APD_DEBUG=1 timeout 400 python bench.py --workload C5 --seqs 300 --steps 1 --warmup 1 --e2e-steps 1 --other-mode-steps 0 --no-cpu --no-parity > gpurun_out/r2d_c5_300_dbg.json 2> gpurun_out/r2d_c5_300_dbg.err; echo "rc=$?"
grep "paths:" gpurun_out/r2d_c5_300_dbg.err | tail -12
timeout 200 python bench.py --seqs 4000 --steps 3 --warmup 2 --no-cpu --no-parity --e2e-steps 1 > gpurun_out/r2d_c3_4000_default.json 2> gpurun_out/r2d_c3_4000_default.err; echo "rc=$?"
APD_LIB_PATH=$PWD/audio_pattern_discovery_b200/libapd_b200.noedge.so timeout 200 python bench.py --seqs 4000 --steps 3 --warmup 2 --no-cpu --no-parity --e2e-steps 1 > gpurun_out/r2d_c3_4000_noedge.json 2> gpurun_out/r2d_c3_4000_noedge.err; echo "rc=$?"
python tools/profile_case.py --workload C2 --n 2000 --mode strict > gpurun_out/r2d_prof_c2_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:dtw_units -c 3 -o gpurun_out/r2d_ncu_c2_strict -f python tools/profile_case.py --workload C2 --n 2000 --mode strict > gpurun_out/r2d_ncu_c2.log 2>&1; echo "ncu rc=$?"
python tools/profile_case.py --workload C3 --n 1000 --mode strict > gpurun_out/r2d_prof_c3_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:dtw_units -c 1 -o gpurun_out/r2d_ncu_c3_strict -f python tools/profile_case.py --workload C3 --n 1000 --mode strict > gpurun_out/r2d_ncu_c3.log 2>&1; echo "ncu rc=$?"
timeout 900 python bench.py --workload C4 --steps 1 --warmup 1 --e2e-steps 1 --other-mode-steps 1 --no-cpu > gpurun_out/r2d_c4_full.json 2> gpurun_out/r2d_c4_full.err; echo "rc=$?"
ls -la gpurun_out | tail -20
