set -x
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2j_pytest_gpu_2gpus.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j_pytest_gpu_2gpus.log
tail -4 gpurun_out/r2j_pytest_gpu_2gpus.log
timeout 300 python bench.py --gpus 2 --single-process --steps 3 --warmup 2 --no-cpu --e2e-steps 1 > gpurun_out/r2j_bench_C3_n2_single_process.json 2> gpurun_out/r2j_sp2.err; echo "rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 2 --no-cpu --e2e-steps 1 --other-mode-steps 1 > gpurun_out/r2j_bench_C3_n2_torchrun.json 2> gpurun_out/r2j_tr2.err; echo "rc=$?"
tail -n 3 gpurun_out/r2j_*.err
