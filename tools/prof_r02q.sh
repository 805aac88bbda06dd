cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for i in 1 2 3 4 5 6 7 8; do
timeout 300 python -m pytest tests/test_gpu_onchip.py -m gpu -q -x > gpurun_out/r02q_run$i.log 2>&1; echo "run $i rc=$?"; grep -E "ierr differs|passed|failed" gpurun_out/r02q_run$i.log | cut -c1-400
done
