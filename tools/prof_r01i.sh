# round 1, ninth batch: persistent double-buffered cw_rc kernel: tests, timing, ncu.
set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_cwrc.py tests/test_gpu_fastkmt.py -x -q > gpurun_out/t_i.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/t_i.log
A="python bench.py --cols 10000 --mechs gas --steps 3 --warmup 3 --spinup 1 --no-e2e --kon-layers 500 --no-cpu-baseline"
timeout 900 $A > gpurun_out/bench_r01i.json 2> gpurun_out/bench_r01i.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_r01i.err
python - <<'P'
import json
d = json.loads(open("gpurun_out/bench_r01i.json").read().strip().splitlines()[-1])
x = d["next_rows"]["cw_rc"]
print("cw_rc", x["ms_per_step"], x["roofline"]["achieved"], x["roofline"]["frac"])
P
B="python bench.py --cols 200 --mechs gas --steps 1 --warmup 1 --spinup 1 --no-cpu-baseline --no-e2e --kon-layers 500"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'cwrc' -c 3 -o gpurun_out/prof_r01i $B > gpurun_out/ncu_r01i.log 2>&1
echo "ncu rc=$?"
ncu -i gpurun_out/prof_r01i.ncu-rep --page raw --csv > gpurun_out/prof_r01i_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_r01i.ncu-rep --page source --csv > gpurun_out/prof_r01i_src.csv 2>/dev/null
rm -f gpurun_out/prof_r01i.ncu-rep
