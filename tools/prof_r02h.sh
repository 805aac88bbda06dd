cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
OC_STRICT=1 MISTRA_KPP_LIB=libmistra_kpp_ocdbgs.so timeout 120 python tools/oc_debug.py gas 0 > gpurun_out/r02h_dbg_gas.txt 2>&1 || { echo "gas debug failed rc=$?"; tail -5 gpurun_out/r02h_dbg_gas.txt; exit 0; }
cut -c1-200 gpurun_out/r02h_dbg_gas.txt | tail -14
OC_STRICT=1 MISTRA_KPP_LIB=libmistra_kpp_ocdbgs.so timeout 120 python tools/oc_debug.py aer 0 > gpurun_out/r02h_dbg_aer.txt 2>&1 || { echo "aer debug failed rc=$?"; tail -5 gpurun_out/r02h_dbg_aer.txt; exit 0; }
cut -c1-200 gpurun_out/r02h_dbg_aer.txt | tail -14
timeout 300 python tools/oc_bench.py gas 300 3 500 > gpurun_out/r02h_oc_gas.txt 2>&1; tail -3 gpurun_out/r02h_oc_gas.txt
timeout 300 python tools/oc_bench.py aer 300 3 500 > gpurun_out/r02h_oc_aer.txt 2>&1; tail -3 gpurun_out/r02h_oc_aer.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r02h_parity.log 2>&1; echo "parity rc=$?"; tail -5 gpurun_out/r02h_parity.log
