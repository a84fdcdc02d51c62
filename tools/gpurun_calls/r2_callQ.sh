set -x
V=$PWD/audio_pattern_discovery_b200/libapd_b200.colpack.so
B="--seqs 4000 --steps 3 --warmup 2 --no-cpu --no-parity --e2e-steps 1 --other-mode-steps 0"
timeout 200 python bench.py $B > gpurun_out/r2q_c3_4000_default.json 2> gpurun_out/r2q_def.err; echo "rc=$?"
APD_LIB_PATH=$V timeout 200 python bench.py $B > gpurun_out/r2q_c3_4000_colpack_speed_only.json 2> gpurun_out/r2q_cp.err; echo "rc=$?"
