set -x
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2ab_pytest_gpu_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2ab_pytest_gpu_final.log
tail -3 gpurun_out/r2ab_pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.build(); g.smoke()"; echo "smoke rc=$?"
timeout 300 python bench.py --seqs 4000 --steps 3 --warmup 3 --no-cpu > gpurun_out/r2ab_bench_c3_4000.json 2> gpurun_out/r2ab.err; echo "rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r2ab_bench_c3_4000.json')); print(round(d['value'],1), d['parity_ok'], d['matrix_checksum_u64'])"
