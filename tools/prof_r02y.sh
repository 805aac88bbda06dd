cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_rates.py -m gpu -q -x > gpurun_out/r02y_rates.log 2>&1; echo "rates tests rc=$?"; tail -3 gpurun_out/r02y_rates.log | cut -c1-300
(time timeout 1500 python bench.py > gpurun_out/r02y_bench_1gpu.json 2> gpurun_out/r02y_bench_1gpu.err); echo "bench rc=$?"; tail -5 gpurun_out/r02y_bench_1gpu.err
(time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02y_bench_ref.json 2> gpurun_out/r02y_bench_ref.err); echo "ref rc=$?"
python - <<'PY'
import json
l = json.loads(open("gpurun_out/r02y_bench_1gpu.json").read().strip().splitlines()[-1])
for k in ("value", "ms_per_step", "e2e", "parity", "tot", "cold_start", "onchip_aer", "latency_1cell", "per_mechanism", "clocks", "gpu_launches"):
    print(k, json.dumps(l.get(k))[:1200])
print("roofline", json.dumps(l["roofline"])[:1500])
print("cpu", json.dumps(l["cpu_baseline"])[:600])
r = json.loads(open("gpurun_out/r02y_bench_ref.json").read().strip().splitlines()[-1])
print("ref", r["value"], r["cpu_baseline"])
PY
