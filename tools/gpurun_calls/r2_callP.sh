set -x
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2p_pytest_gpu_2gpus.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2p_pytest_gpu_2gpus.log
tail -5 gpurun_out/r2p_pytest_gpu_2gpus.log
