set -x
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_gpu_group.py tests/test_gpu_multi.py -m gpu -q > gpurun_out/r2h_pytest_gpu_8gpus.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_pytest_gpu_8gpus.log
tail -4 gpurun_out/r2h_pytest_gpu_8gpus.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu > gpurun_out/r2h_bench_C3_n8_torchrun.json 2> gpurun_out/r2h_bench_C3_n8_torchrun.err; echo "rc=$?"
timeout 400 python bench.py --gpus 8 --single-process --steps 5 --warmup 3 --no-cpu > gpurun_out/r2h_bench_C3_n8_single_process.json 2> gpurun_out/r2h_bench_C3_n8_single_process.err; echo "rc=$?"
timeout 500 python bench.py --gpus 8 --single-process --workload C4 --steps 2 --warmup 2 --e2e-steps 1 --no-cpu > gpurun_out/r2h_bench_C4_n8_single_process.json 2> gpurun_out/r2h_bench_C4_n8_single_process.err; echo "rc=$?"
timeout 400 python bench.py --gpus 8 --single-process --workload C5 --steps 2 --warmup 2 --e2e-steps 1 --no-cpu > gpurun_out/r2h_bench_C5_n8_single_process.json 2> gpurun_out/r2h_bench_C5_n8_single_process.err; echo "rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 4 --steps 3 --warmup 3 --no-cpu --other-mode-steps 0 > gpurun_out/r2h_bench_C3_n4_torchrun.json 2> gpurun_out/r2h_bench_C3_n4_torchrun.err; echo "rc=$?"
tail -n 3 gpurun_out/r2h_*.err
