cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --no-bins > gpurun_out/r02_bench_8gpu_weak.json 2> gpurun_out/r02_bench_8gpu_weak.err; echo "weak rc=$?"; tail -3 gpurun_out/r02_bench_8gpu_weak.err
python - <<'PY'
import json
l = json.loads(open("gpurun_out/r02_bench_8gpu_weak.json").read().strip().splitlines()[-1])
print("value", l["value"], "e2e", l["e2e"]["value"], "h2d", l["e2e"]["h2d_bytes_per_step"], "rconst path", l["e2e"]["rconst_path"]["value"], "n_gpus", l["n_gpus"])
PY
