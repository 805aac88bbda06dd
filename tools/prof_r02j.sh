cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tools/oc_bench.py aer 300 3 500 > gpurun_out/r02j_oc_aer.txt 2>&1; tail -3 gpurun_out/r02j_oc_aer.txt
timeout 300 python tools/oc_bench.py gas 300 3 500 > gpurun_out/r02j_oc_gas.txt 2>&1; tail -3 gpurun_out/r02j_oc_gas.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r02j_parity.log 2>&1; echo "parity rc=$?"; tail -5 gpurun_out/r02j_parity.log
