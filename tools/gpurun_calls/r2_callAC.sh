set -x
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_group.py -m gpu -q > gpurun_out/r2ac_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2ac_pytest.log
tail -3 gpurun_out/r2ac_pytest.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu --e2e-steps 3 > gpurun_out/r2ac_bench_C3_n2_torchrun.json 2> gpurun_out/r2ac_tr2.err; echo "rc=$?"
tail -n 3 gpurun_out/r2ac_tr2.err
