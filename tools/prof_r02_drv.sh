cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_driver.py tests/test_gpu_sed.py -m gpu -q -x > gpurun_out/r02_drv_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r02_drv_tests.log
