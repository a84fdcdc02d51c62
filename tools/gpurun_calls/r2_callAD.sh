set -x
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29591 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu --e2e-steps 3 > gpurun_out/r2ad_bench_C3_n8_torchrun.json 2> gpurun_out/r2ad_tr8.err; echo "rc=$?"
tail -n 2 gpurun_out/r2ad_tr8.err
