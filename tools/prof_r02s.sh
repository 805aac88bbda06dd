cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python bench.py --cols 400 --steps 2 --warmup 1 --tot-cells 9800 --kon-layers 500 --bins-layers 1480 > gpurun_out/r02s_bench_small.json 2> gpurun_out/r02s_bench_small.err; echo "bench rc=$?"; tail -5 gpurun_out/r02s_bench_small.err; python - <<'PY'
import json
l = json.loads(open("gpurun_out/r02s_bench_small.json").read().strip().splitlines()[-1])
for k in ("value", "e2e", "parity", "tot", "cold_start", "onchip_aer", "latency_1cell", "per_mechanism"):
    print(k, json.dumps(l.get(k))[:900])
PY
timeout 600 python tools/spinup_check.py > gpurun_out/r02s_spinup.txt 2>&1; cat gpurun_out/r02s_spinup.txt
