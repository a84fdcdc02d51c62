set -x
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2z_bench_C3_n2_driver_args.json 2> gpurun_out/r2z_tr2.err; echo "rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29572 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2z_reference_n2.json 2> gpurun_out/r2z_ref.err; echo "rc=$?"
wc -l gpurun_out/r2z_bench_C3_n2_driver_args.json gpurun_out/r2z_reference_n2.json
