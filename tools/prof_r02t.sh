cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python tools/variant_sweep.py > gpurun_out/r02t_sweep.txt 2>&1; cat gpurun_out/r02t_sweep.txt
nfail=0
for i in $(seq 1 24); do
timeout 300 python -m pytest tests/test_gpu_zz_onchip.py -m gpu -q -x > gpurun_out/r02t_run.log 2>&1 || { nfail=$((nfail+1)); cp gpurun_out/r02t_run.log gpurun_out/r02t_fail$nfail.log; grep -E "ierr differs|Error|assert" gpurun_out/r02t_run.log | head -5 | cut -c1-600; }
done
echo "module failures: $nfail of 24"
