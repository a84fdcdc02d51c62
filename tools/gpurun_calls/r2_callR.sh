set -x
timeout 600 python -m pytest tests/test_gpu_group.py tests/test_gpu_parity.py -m gpu -q -k "group or hybrid or wide" > gpurun_out/r2r_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2r_pytest.log
tail -15 gpurun_out/r2r_pytest.log
