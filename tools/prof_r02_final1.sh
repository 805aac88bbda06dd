cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -12 gpurun_out/r02_smoke.log
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r02_gpu_suite.log 2>&1; echo "gpu suite rc=$?"; tail -4 gpurun_out/r02_gpu_suite.log
