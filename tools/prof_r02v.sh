cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r02v_gpu_suite.log 2>&1; echo "gpu suite rc=$?"; tail -6 gpurun_out/r02v_gpu_suite.log
timeout 300 python tools/oc_stress.py 20 > gpurun_out/r02v_stress.txt 2>&1; tail -9 gpurun_out/r02v_stress.txt
timeout 300 python tools/oc_bench.py aer 300 3 500 > gpurun_out/r02v_oc_aer.txt 2>&1; tail -2 gpurun_out/r02v_oc_aer.txt
