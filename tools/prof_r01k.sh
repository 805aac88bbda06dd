# round 1, eleventh batch: device-chain test, difc / difp with the zero-numerator shortcut (dense and sparse inputs).
set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_chain.py tests/test_gpu_difc.py -x -q > gpurun_out/t_k.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/t_k.log
timeout 300 python tools/difc_sweep.py | tee gpurun_out/difc_sweep5.txt
