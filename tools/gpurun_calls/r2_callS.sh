set -x
V=$PWD/audio_pattern_discovery_b200/libapd_b200.xup.so
B="--seqs 4000 --steps 3 --warmup 2 --no-cpu --e2e-steps 1 --other-mode-steps 1"
APD_WIDE=1 timeout 200 python bench.py $B > gpurun_out/r2s_c3_4000_wide.json 2> gpurun_out/r2s_w.err; echo "rc=$?"
APD_WIDE=1 APD_LIB_PATH=$V timeout 200 python bench.py $B > gpurun_out/r2s_c3_4000_wide_xup.json 2> gpurun_out/r2s_wx.err; echo "rc=$?"
