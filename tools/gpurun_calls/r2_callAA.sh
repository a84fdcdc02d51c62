set -x
B="--seqs 4000 --steps 3 --warmup 2 --no-cpu --e2e-steps 1 --other-mode-steps 1"
for v in look3 look1; do
  APD_LIB_PATH=$PWD/audio_pattern_discovery_b200/libapd_b200.$v.so timeout 200 python bench.py $B > gpurun_out/r2aa_c3_4000_$v.json 2> gpurun_out/r2aa_$v.err; echo "rc=$?"
done
timeout 200 python bench.py $B > gpurun_out/r2aa_c3_4000_default.json 2> gpurun_out/r2aa_d.err; echo "rc=$?"
