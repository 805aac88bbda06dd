cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
MISTRA_KPP_LIB=libmistra_kpp_ocph.so timeout 300 python tools/oc_stall.py > gpurun_out/r02u_stall.txt 2>&1; cat gpurun_out/r02u_stall.txt
