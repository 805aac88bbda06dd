cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
OC_STRICT=1 MISTRA_KPP_LIB=libmistra_kpp_ocdbgs.so timeout 600 python tools/oc_debug.py gas 0 > gpurun_out/r02c_dbg_gas.txt 2>&1
OC_STRICT=1 MISTRA_KPP_LIB=libmistra_kpp_ocdbgs.so timeout 600 python tools/oc_debug.py aer 0 > gpurun_out/r02c_dbg_aer.txt 2>&1
