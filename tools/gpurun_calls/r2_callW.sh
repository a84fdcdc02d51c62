set -x
timeout 600 python bench.py > gpurun_out/r2w_bench_C3_n1_final.json 2> gpurun_out/r2w.err; echo "rc=$?"
tail -n 2 gpurun_out/r2w.err
