set -x
python tools/host_copy_probe.py > gpurun_out/r2a_host_copy_probe.json 2> gpurun_out/r2a_host_copy_probe.err
APD_DEBUG=1 python bench.py --workload C2 --steps 5 --warmup 3 --no-cpu > gpurun_out/r2a_c2.json 2> gpurun_out/r2a_c2.err
APD_DEBUG=1 python bench.py --workload C1ref --steps 5 --warmup 3 --no-cpu > gpurun_out/r2a_c1ref.json 2> gpurun_out/r2a_c1ref.err
APD_DEBUG=1 python bench.py --workload C5 --seqs 500 --steps 1 --warmup 1 --e2e-steps 1 --other-mode-steps 1 --no-cpu > gpurun_out/r2a_c5_500.json 2> gpurun_out/r2a_c5_500.err
APD_DEBUG=1 python bench.py --workload C4 --seqs 5000 --steps 1 --warmup 1 --e2e-steps 1 --other-mode-steps 1 --no-cpu > gpurun_out/r2a_c4_5000.json 2> gpurun_out/r2a_c4_5000.err
nvidia-smi -L; nproc; free -g | head -2
