cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_onchip.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r02m_tests.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/r02m_tests.log
