cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 > gpurun_out/r02_bench_2gpu_weak.json 2> gpurun_out/r02_bench_2gpu_weak.err; echo "weak rc=$?"; tail -3 gpurun_out/r02_bench_2gpu_weak.err
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --scaling strong --no-bins > gpurun_out/r02_bench_2gpu_strong.json 2> gpurun_out/r02_bench_2gpu_strong.err; echo "strong rc=$?"
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/r02_multi_2gpu.log 2>&1; tail -2 gpurun_out/r02_multi_2gpu.log
python - <<'PY'
import json
for f in ("weak", "strong"):
    try:
        l = json.loads(open("gpurun_out/r02_bench_2gpu_%s.json" % f).read().strip().splitlines()[-1])
        print(f, "value", l["value"], "e2e", l["e2e"]["value"], "rconst path", l["e2e"]["rconst_path"]["value"], "scaling", l["scaling"], "n_gpus", l["n_gpus"], "cells/gpu", l["config"]["cells_per_gpu"])
    except Exception as e:
        print(f, "failed", e)
PY
