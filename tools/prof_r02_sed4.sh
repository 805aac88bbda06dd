cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for lib in libmistra_kpp.so libmistra_kpp_sedA.so libmistra_kpp_sedB.so libmistra_kpp_sedC.so; do echo "== $lib"; MISTRA_KPP_LIB=$lib timeout 300 python tools/sed_bench.py 512; done 2>&1 | tee gpurun_out/r02_sed_minb_sweep.txt
