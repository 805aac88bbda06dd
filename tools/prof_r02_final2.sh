cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
(time timeout 1500 python bench.py > gpurun_out/r02_final_bench_1gpu.json 2> gpurun_out/r02_final_bench_1gpu.err); echo "bench rc=$?"; tail -4 gpurun_out/r02_final_bench_1gpu.err
python - <<'PY'
import json
l = json.loads(open("gpurun_out/r02_final_bench_1gpu.json").read().strip().splitlines()[-1])
for k in ("value", "ms_per_step", "gpu_launches", "per_mechanism", "clocks"):
    print(k, json.dumps(l.get(k))[:400])
e = l["e2e"]; print("e2e", e["value"], e["ms_per_step"], e["h2d_bytes_per_step"], "rconst path", e["rconst_path"]["value"], "pageable", e["pageable"]["value"])
r = l["roofline"]; print("roofline hbm frac", r["frac"], "fp64 frac", r["fp64"]["frac"], "kernel_ms", r["kernel_ms"])
print("parity", l["parity"]["ok"], "tot", l["tot"]["cells_per_s"], "onchip", l["onchip_aer"]["cells_per_s"], l["onchip_aer"]["cell_per_thread_kernel_same_cells"], "lat", l["latency_1cell"]["gas"]["b1_us_per_call"], l["latency_1cell"]["aer"]["b1_us_per_call"])
PY
