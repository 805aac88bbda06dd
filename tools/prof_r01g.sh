# round 1, seventh batch: difc batch depth sweep (U = 8 | 16), fast_k_mt with staged layer scalars.
set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_difc.py tests/test_gpu_fastkmt.py -x -q > gpurun_out/t_g.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/t_g.log
for u in 8 16; do for r in 16 5 4 3; do MISTRA_DIFC_U=$u MISTRA_DIFC_CTAS_PER_SM=$r timeout 300 python tools/difc_sweep.py | sed "s/^/U=$u /"; done; done 2>&1 | grep CTAs | tee gpurun_out/difc_sweep3.txt
MISTRA_DIFC_U=16 timeout 300 python -m pytest tests/test_gpu_difc.py -x -q 2>&1 | tail -2
A="python bench.py --cols 10000 --mechs gas --steps 3 --warmup 3 --spinup 1 --no-e2e --kon-layers 500 --no-cpu-baseline"
timeout 900 $A > gpurun_out/bench_r01g.json 2> gpurun_out/bench_r01g.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_r01g.err
python - <<'P'
import json
d = json.loads(open("gpurun_out/bench_r01g.json").read().strip().splitlines()[-1])
x = d["next_rows"]["fast_k_mt"]
print("fast_k_mt", {k: (v["ms_per_step"], v["fp64"]["divisions_per_s"]) for k, v in x["cases"].items()})
P
B="python bench.py --cols 200 --mechs gas --steps 1 --warmup 1 --spinup 1 --no-cpu-baseline --no-e2e --kon-layers 500 --bins-layers 2960"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'fastkmt' -c 4 -o gpurun_out/prof_r01g $B > gpurun_out/ncu_r01g.log 2>&1
echo "ncu rc=$?"
ncu -i gpurun_out/prof_r01g.ncu-rep --page raw --csv > gpurun_out/prof_r01g_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_r01g.ncu-rep --page source --csv > gpurun_out/prof_r01g_src.csv 2>/dev/null
rm -f gpurun_out/prof_r01g.ncu-rep
