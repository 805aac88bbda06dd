cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_gpu_zz_onchip.py -m gpu -q -x -k handoff ) > gpurun_out/r02_handoff_tests.log 2>&1; echo "tests rc=$?"; tail -12 gpurun_out/r02_handoff_tests.log
bash tools/prof_r02_handoff2.sh
