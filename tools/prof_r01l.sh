# round 1, twelfth batch: smoke, drive leg of the bench.
set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
timeout 900 python bench.py --cols 10000 --mechs gas --steps 3 --warmup 3 --spinup 1 --no-e2e --kon-layers 500 --no-cpu-baseline > gpurun_out/bench_r01l.json 2> gpurun_out/bench_r01l.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_r01l.err
python - <<'P'
import json
d = json.loads(open("gpurun_out/bench_r01l.json").read().strip().splitlines()[-1])
print(json.dumps(d["next_rows"]["drive"])[:900])
P
