cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_b1.py tests/test_gpu_multi.py -m gpu -q > gpurun_out/r02r_tests.log 2>&1; echo "tests rc=$?"; tail -12 gpurun_out/r02r_tests.log
