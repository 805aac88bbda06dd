# round 1, eighth batch: fast_k_mt with the two-pass compaction; ncu full of fastkmt_kernel, difc_solve_kernel<1> and difp_fsum_kernel.
set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_fastkmt.py -x -q > gpurun_out/t_h.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/t_h.log
A="python bench.py --cols 10000 --mechs gas --steps 3 --warmup 3 --spinup 1 --no-e2e --kon-layers 500 --no-cpu-baseline"
timeout 900 $A > gpurun_out/bench_r01h.json 2> gpurun_out/bench_r01h.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_r01h.err
python - <<'P'
import json
d = json.loads(open("gpurun_out/bench_r01h.json").read().strip().splitlines()[-1])
x = d["next_rows"]["fast_k_mt"]
print("fast_k_mt", {k: (v["ms_per_step"], v["fp64"]["divisions_per_s"]) for k, v in x["cases"].items()})
P
B="python bench.py --cols 200 --mechs gas --steps 1 --warmup 1 --spinup 1 --no-cpu-baseline --no-e2e --kon-layers 500 --bins-layers 2960"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'fastkmt' -c 4 -o gpurun_out/prof_r01h $B > gpurun_out/ncu_r01h.log 2>&1
echo "ncu rc=$?"
ncu -i gpurun_out/prof_r01h.ncu-rep --page raw --csv > gpurun_out/prof_r01h_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_r01h.ncu-rep --page source --csv > gpurun_out/prof_r01h_src.csv 2>/dev/null
rm -f gpurun_out/prof_r01h.ncu-rep
timeout 600 ncu --set full --clock-control none -k regex:'difc_solve_kernel<1|difp_fsum' -c 4 -o gpurun_out/prof_r01h2 python tools/difc_sweep.py > gpurun_out/ncu_r01h2.log 2>&1
echo "ncu2 rc=$?"
ncu -i gpurun_out/prof_r01h2.ncu-rep --page raw --csv > gpurun_out/prof_r01h2_raw.csv 2>/dev/null
rm -f gpurun_out/prof_r01h2.ncu-rep
