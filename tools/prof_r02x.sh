cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rates.py -m gpu -q -x > gpurun_out/r02x_rates.log 2>&1; echo "rates tests rc=$?"; tail -15 gpurun_out/r02x_rates.log | cut -c1-300
timeout 900 python bench.py --cols 400 --steps 2 --warmup 1 --tot-cells 9800 --no-bins > gpurun_out/r02x_bench_small.json 2> gpurun_out/r02x_bench_small.err; echo "bench rc=$?"; tail -5 gpurun_out/r02x_bench_small.err; python - <<'PY'
import json
l = json.loads(open("gpurun_out/r02x_bench_small.json").read().strip().splitlines()[-1])
for k in ("value", "e2e", "parity", "latency_1cell"):
    print(k, json.dumps(l.get(k))[:1500])
PY
