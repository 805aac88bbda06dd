cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r02w_gpu_suite.log 2>&1; echo "gpu suite rc=$?"; tail -4 gpurun_out/r02w_gpu_suite.log
nfail=0
for i in $(seq 1 12); do
timeout 300 python -m pytest tests/test_gpu_zz_onchip.py -m gpu -q -x > gpurun_out/r02w_run.log 2>&1 || { nfail=$((nfail+1)); cp gpurun_out/r02w_run.log gpurun_out/r02w_fail$nfail.log; }
done
echo "on-chip module failures: $nfail of 12"
