cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tools/oc_bench.py aer 300 3 500 > gpurun_out/r02l_oc_aer.txt 2>&1; tail -3 gpurun_out/r02l_oc_aer.txt
timeout 300 python tools/oc_bench.py gas 300 3 500 > gpurun_out/r02l_oc_gas.txt 2>&1; tail -3 gpurun_out/r02l_oc_gas.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r02l_parity.log 2>&1; echo "parity rc=$?"; tail -5 gpurun_out/r02l_parity.log
OC_PHASES=1 MISTRA_KPP_LIB=libmistra_kpp_ocph.so timeout 100 python tools/oc_bench.py aer 300 3 0 > gpurun_out/r02l_ph_aer.txt 2>&1; tail -14 gpurun_out/r02l_ph_aer.txt
