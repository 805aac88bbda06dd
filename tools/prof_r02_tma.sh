cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cwrc.py tests/test_gpu_fastkmt.py tests/test_gpu_chain.py tests/test_gpu_fullsize_properties.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r02_tma_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r02_tma_tests.log
timeout 1200 python bench.py --no-e2e --no-extras --no-cpu-baseline > gpurun_out/r02_tma_bench.json 2> gpurun_out/r02_tma_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_tma_bench.err
python - <<'PY'
import json
l = json.loads(open("gpurun_out/r02_tma_bench.json").read().strip().splitlines()[-1])
print("value", l["value"], "ms", l["ms_per_step"], "launches", l["gpu_launches"], "per mech", l["per_mechanism"])
nr = l["next_rows"]
for k in ("cw_rc", "fast_k_mt"):
    if k in nr: print(k, json.dumps(nr[k])[:700])
print("parity", l["parity"]["ok"])
PY
