set -x
B="--no-cpu --no-parity --e2e-steps 1 --other-mode-steps 1"
for w in C1ref C2; do
  timeout 200 python bench.py --workload $w --steps 5 --warmup 3 $B > gpurun_out/r2m_${w}_carveout.json 2> gpurun_out/r2m_${w}_carveout.err; echo "rc=$?"
  APD_CARVEOUT=0 timeout 200 python bench.py --workload $w --steps 5 --warmup 3 $B > gpurun_out/r2m_${w}_nocarveout.json 2> gpurun_out/r2m_${w}_nocarveout.err; echo "rc=$?"
done
timeout 300 python bench.py --workload C4 --seqs 5000 --steps 2 --warmup 1 $B > gpurun_out/r2m_C4_5000_carveout.json 2> gpurun_out/r2m_C4_carveout.err; echo "rc=$?"
APD_CARVEOUT=0 timeout 300 python bench.py --workload C4 --seqs 5000 --steps 2 --warmup 1 $B > gpurun_out/r2m_C4_5000_nocarveout.json 2> gpurun_out/r2m_C4_nocarveout.err; echo "rc=$?"
