cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
OC_PHASES=1 MISTRA_KPP_LIB=libmistra_kpp_ocph.so timeout 100 python tools/oc_bench.py aer 300 3 0 > gpurun_out/r02k_ph_aer.txt 2>&1; tail -16 gpurun_out/r02k_ph_aer.txt
OC_PHASES=1 MISTRA_KPP_LIB=libmistra_kpp_ocph.so timeout 100 python tools/oc_bench.py gas 300 3 0 > gpurun_out/r02k_ph_gas.txt 2>&1; tail -16 gpurun_out/r02k_ph_gas.txt
