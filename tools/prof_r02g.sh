cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
OC_PHASES=1 MISTRA_KPP_LIB=libmistra_kpp_ocph.so timeout 200 compute-sanitizer --tool memcheck --print-limit 5 python tools/oc_bench.py aer 1 1 0 > gpurun_out/r02g_san_aer.txt 2>&1; tail -40 gpurun_out/r02g_san_aer.txt
