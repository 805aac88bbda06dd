cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
MISTRA_KPP_LIB=libmistra_kpp_ocdbg.so timeout 600 python tools/oc_debug.py gas 0 > gpurun_out/r02b_dbg_gas.txt 2>&1
MISTRA_KPP_LIB=libmistra_kpp_ocdbg.so timeout 600 python tools/oc_debug.py aer 0 > gpurun_out/r02b_dbg_aer.txt 2>&1
tail -5 gpurun_out/r02b_dbg_gas.txt
