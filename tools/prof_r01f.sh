# round 1, sixth batch: batched (software-pipelined) difc / difp sweeps, fast_k_mt without the division slow path.
set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_difc.py tests/test_gpu_fastkmt.py -x -q > gpurun_out/t_f.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/t_f.log
for r in 16 6 4 3 2; do MISTRA_DIFC_CTAS_PER_SM=$r timeout 300 python tools/difc_sweep.py; done 2>&1 | grep CTAs | tee gpurun_out/difc_sweep2.txt
A="python bench.py --cols 10000 --mechs gas --steps 3 --warmup 3 --spinup 1 --no-e2e --kon-layers 500"
timeout 900 $A > gpurun_out/bench_r01f.json 2> gpurun_out/bench_r01f.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_r01f.err
python - <<'P'
import json
d = json.loads(open("gpurun_out/bench_r01f.json").read().strip().splitlines()[-1])
for k in ("fast_k_mt", "difc", "difp"):
    x = d.get("next_rows", {}).get(k) or {}
    print(k, x.get("value"), x.get("ms_per_step"), json.dumps(x.get("roofline"))[:300], json.dumps(x.get("cases"))[:900], json.dumps(x.get("cpu_baseline"))[:200])
P
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'difc_solve|difp_' -c 6 -o gpurun_out/prof_r01f python tools/difc_sweep.py > gpurun_out/ncu_r01f.log 2>&1
echo "ncu rc=$?"
ncu -i gpurun_out/prof_r01f.ncu-rep --page raw --csv > gpurun_out/prof_r01f_raw.csv 2>/dev/null
rm -f gpurun_out/prof_r01f.ncu-rep
