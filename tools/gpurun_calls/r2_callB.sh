set -x
nvidia-smi -L
nvidia-smi topo -m 2>/dev/null | head -12
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest_gpu.log
tail -5 gpurun_out/r2b_pytest_gpu.log
timeout 300 python bench.py --gpus 2 --single-process --seqs 4000 --steps 3 --warmup 2 --no-cpu > gpurun_out/r2b_sp2_c3_4000.json 2> gpurun_out/r2b_sp2_c3_4000.err; echo "rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --seqs 4000 --steps 3 --warmup 2 --no-cpu --other-mode-steps 1 > gpurun_out/r2b_tr2_c3_4000.json 2> gpurun_out/r2b_tr2_c3_4000.err; echo "rc=$?"
timeout 300 python bench.py --gpus 1 --seqs 4000 --steps 3 --warmup 2 --no-cpu > gpurun_out/r2b_n1_c3_4000.json 2> gpurun_out/r2b_n1_c3_4000.err; echo "rc=$?"
timeout 200 python bench.py --gpus 1 --workload C2 --steps 5 --warmup 3 --no-cpu > gpurun_out/r2b_c2.json 2> gpurun_out/r2b_c2.err; echo "rc=$?"
APD_SERIAL_CLASSES=1 timeout 200 python bench.py --gpus 1 --workload C2 --steps 5 --warmup 3 --no-cpu --no-parity > gpurun_out/r2b_c2_serial.json 2> gpurun_out/r2b_c2_serial.err; echo "rc=$?"
tail -3 gpurun_out/*.err
