set -x
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29601 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2af_bench_C3_n8_driver_args.json 2> gpurun_out/r2af.err; echo "rc=$?"
wc -l gpurun_out/r2af_bench_C3_n8_driver_args.json
