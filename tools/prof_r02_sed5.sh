cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"sedp_work|sedl_kernel|sedp_scan" -c 5 -o gpurun_out/r02_sed_inplace --force-overwrite python tools/sed_bench.py 256 > gpurun_out/r02_sed_ncu3.log 2>&1; tail -3 gpurun_out/r02_sed_ncu3.log
