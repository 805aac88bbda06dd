cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q -x ) > gpurun_out/r03_gpu_suite.log 2>&1; echo "suite rc=$?"; tail -5 gpurun_out/r03_gpu_suite.log
( time timeout 1500 python bench.py ) > gpurun_out/r03_bench_1gpu.json 2> gpurun_out/r03_bench_1gpu.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r03_bench_1gpu.json; tail -5 gpurun_out/r03_bench_1gpu.err
( time timeout 600 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r03_smoke.log 2>&1; echo "smoke rc=$?"; tail -8 gpurun_out/r03_smoke.log
