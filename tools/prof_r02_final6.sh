cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 200 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r06_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/r06_smoke.log
( timeout 300 python -m pytest tests/test_gpu_zz_onchip.py tests/test_gpu_sed.py tests/test_gpu_driver.py -m gpu -q -x ) > gpurun_out/r06_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r06_tests.log
