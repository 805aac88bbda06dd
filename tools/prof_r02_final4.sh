cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/r04_gpu_suite.log 2>&1; echo "suite rc=$?"; tail -5 gpurun_out/r04_gpu_suite.log
( time timeout 1500 python bench.py ) > gpurun_out/r04_bench_1gpu.json 2> gpurun_out/r04_bench_1gpu.err; echo "bench rc=$?"; tail -4 gpurun_out/r04_bench_1gpu.err
( time timeout 900 python bench.py --impl reference ) > gpurun_out/r04_bench_ref.json 2> gpurun_out/r04_bench_ref.err; echo "ref rc=$?"; tail -c 400 gpurun_out/r04_bench_ref.json
( time timeout 600 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r04_smoke.log 2>&1; echo "smoke rc=$?"; tail -6 gpurun_out/r04_smoke.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r04_launches.csv python bench.py --cols 500 --steps 2 --warmup 1 --spinup 2 --no-cpu-baseline --no-e2e --no-extras --kon-layers 2000 --bins-layers 2960 > gpurun_out/r04_ncu.log 2>&1; echo "ncu rc=$?"
