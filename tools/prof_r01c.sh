# round 1, third batch: GPU tests of fast_k_mt, whole GPU suite, small bench (next rows), ncu of fastkmt_kernel.
# Run under gpurun: bash tools/prof_r01c.sh
set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_fastkmt.py -x -q > gpurun_out/t_fastkmt.log 2>&1; echo "fastkmt tests rc=$?"; tail -5 gpurun_out/t_fastkmt.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/t_gpu.log 2>&1; echo "gpu suite rc=$?"; tail -5 gpurun_out/t_gpu.log
A="python bench.py --cols 500 --steps 3 --warmup 3 --spinup 2 --no-e2e --kon-layers 2000"
timeout 900 $A > gpurun_out/bench_small_r01c.json 2> gpurun_out/bench_small_r01c.err; echo "bench rc=$?"; tail -c 6000 gpurun_out/bench_small_r01c.json
B="python bench.py --cols 200 --steps 1 --warmup 1 --spinup 1 --no-cpu-baseline --no-e2e --kon-layers 500 --bins-layers 2960"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'fastkmt_kernel' -c 4 -o gpurun_out/prof_fastkmt $B > gpurun_out/ncu_fastkmt.log 2>&1
echo "fastkmt full rc=$?"
ncu -i gpurun_out/prof_fastkmt.ncu-rep --page raw --csv > gpurun_out/prof_fastkmt_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_fastkmt.ncu-rep --page source --csv > gpurun_out/prof_fastkmt_src.csv 2>/dev/null
ls -la gpurun_out/ | tail -12
