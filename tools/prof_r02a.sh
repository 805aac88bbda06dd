set -x
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
./build_tools/ubench_fp64 > gpurun_out/r02a_ubench.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/r02a_parity.log 2>&1; echo "parity rc=$?"; tail -15 gpurun_out/r02a_parity.log
timeout 600 python tools/oc_bench.py gas 300 3 500 > gpurun_out/r02a_oc_gas.txt 2>&1; tail -3 gpurun_out/r02a_oc_gas.txt
timeout 600 python tools/oc_bench.py aer 300 3 500 > gpurun_out/r02a_oc_aer.txt 2>&1; tail -3 gpurun_out/r02a_oc_aer.txt
MISTRA_KPP_ONCHIP=0 timeout 600 python tools/oc_bench.py aer 300 3 0 > gpurun_out/r02a_old_aer.txt 2>&1; tail -2 gpurun_out/r02a_old_aer.txt
MISTRA_KPP_ONCHIP=0 timeout 600 python tools/oc_bench.py gas 300 3 0 > gpurun_out/r02a_old_gas.txt 2>&1; tail -2 gpurun_out/r02a_old_gas.txt
