set -x
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2ae_pytest_gpu_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2ae_pytest_gpu_final.log
tail -3 gpurun_out/r2ae_pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.build(); g.smoke()"; echo "smoke rc=$?"
