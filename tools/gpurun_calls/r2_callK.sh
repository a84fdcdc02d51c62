set -x
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2k_smoke.log 2>&1; echo "smoke rc=$?"
python tools/profile_case.py --workload C1ref --n 200 --mode strict --reps 3 > gpurun_out/r2k_prof_c1ref_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:dtw_units -c 2 -o gpurun_out/r2k_ncu_c1ref_strict -f python tools/profile_case.py --workload C1ref --n 200 --mode strict > gpurun_out/r2k_ncu_c1ref.log 2>&1; echo "ncu rc=$?"
cat gpurun_out/r2k_prof_c1ref_plain.log
