cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_onchip.py -m gpu -q -k "gas_cold" > gpurun_out/r02o_single.log 2>&1; echo "single rc=$?"; tail -3 gpurun_out/r02o_single.log
timeout 900 python -m pytest tests/test_gpu_onchip.py -m gpu -q > gpurun_out/r02o_all.log 2>&1; echo "all rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/r02o_all.log
