# round 1, final batch: whole GPU suite, default 1-GPU bench line (profiles/), ncu of the drive kernels.
set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/t_m.log 2>&1; echo "gpu suite rc=$?"; tail -15 gpurun_out/t_m.log
timeout 1500 python bench.py > gpurun_out/bench_default_r01d.json 2> gpurun_out/bench_default_r01d.err; echo "bench rc=$?"; tail -c 400 gpurun_out/bench_default_r01d.err
