set -x
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu > gpurun_out/r2l_bench_C3_n8_torchrun.json 2> gpurun_out/r2l_bench_C3_n8_torchrun.err; echo "rc=$?"
timeout 400 python bench.py --gpus 8 --single-process --steps 5 --warmup 3 --no-cpu > gpurun_out/r2l_bench_C3_n8_single_process.json 2> gpurun_out/r2l_bench_C3_n8_single_process.err; echo "rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 4 --steps 3 --warmup 3 --no-cpu --other-mode-steps 0 > gpurun_out/r2l_bench_C3_n4_torchrun.json 2> gpurun_out/r2l_bench_C3_n4_torchrun.err; echo "rc=$?"
tail -n 2 gpurun_out/r2l_*.err
