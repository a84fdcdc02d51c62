set -x
V=$PWD/audio_pattern_discovery_b200/libapd_b200.nocompact.so
B="--workload C2 --steps 5 --warmup 3 --no-cpu --no-parity --e2e-steps 1 --other-mode-steps 0"
timeout 200 python bench.py $B > gpurun_out/r2v_c2_default.json 2> gpurun_out/r2v_d.err; echo "rc=$?"
APD_LIB_PATH=$V timeout 200 python bench.py $B > gpurun_out/r2v_c2_nocompact.json 2> gpurun_out/r2v_n.err; echo "rc=$?"
