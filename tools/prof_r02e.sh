cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
OC_PHASES=1 MISTRA_KPP_LIB=libmistra_kpp_ocph.so timeout 600 python tools/oc_bench.py aer 300 3 0 > gpurun_out/r02e_ph_aer.txt 2>&1; cat gpurun_out/r02e_ph_aer.txt
OC_PHASES=1 MISTRA_KPP_LIB=libmistra_kpp_ocph.so timeout 600 python tools/oc_bench.py gas 300 3 0 > gpurun_out/r02e_ph_gas.txt 2>&1; cat gpurun_out/r02e_ph_gas.txt
timeout 600 python tools/oc_bench.py aer 40 1 0 > gpurun_out/r02e_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ros3_onchip_a -s 2 -c 1 -o gpurun_out/r02e_aer python tools/oc_bench.py aer 40 1 0 > gpurun_out/r02e_ncu.log 2>&1
tail -3 gpurun_out/r02e_ncu.log
