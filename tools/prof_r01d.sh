# round 1, fourth batch: GPU tests of fast_k_mt (compacted lists) and difc, small bench (next rows), ncu of both.
# Run under gpurun: bash tools/prof_r01d.sh
set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_fastkmt.py tests/test_gpu_difc.py -x -q > gpurun_out/t_new.log 2>&1; echo "new tests rc=$?"; tail -15 gpurun_out/t_new.log
A="python bench.py --cols 10000 --mechs gas --steps 3 --warmup 3 --spinup 1 --no-e2e --kon-layers 500"
timeout 900 $A > gpurun_out/bench_r01d.json 2> gpurun_out/bench_r01d.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_r01d.err
python - <<'P'
import json
d = json.loads(open("gpurun_out/bench_r01d.json").read().strip().splitlines()[-1])
for k in ("fast_k_mt", "difc", "cw_rc"):
    print(k, json.dumps(d.get("next_rows", {}).get(k), indent=None)[:2500])
P
B="python bench.py --cols 2000 --mechs gas --steps 1 --warmup 1 --spinup 1 --no-cpu-baseline --no-e2e --kon-layers 500 --bins-layers 2960"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'fastkmt_kernel|difc_' -c 14 -o gpurun_out/prof_r01d $B > gpurun_out/ncu_r01d.log 2>&1
echo "ncu rc=$?"
ncu -i gpurun_out/prof_r01d.ncu-rep --page raw --csv > gpurun_out/prof_r01d_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_r01d.ncu-rep --page source --csv > gpurun_out/prof_r01d_src.csv 2>/dev/null
rm -f gpurun_out/prof_r01d.ncu-rep
ls -la gpurun_out/ | tail -8
