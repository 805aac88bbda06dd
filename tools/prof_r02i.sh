cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python tools/oc_bench.py aer 40 1 0 > gpurun_out/r02i_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ros3_onchip_a -s 2 -c 1 -o gpurun_out/r02i_aer python tools/oc_bench.py aer 40 1 0 > gpurun_out/r02i_ncu.log 2>&1
tail -3 gpurun_out/r02i_ncu.log
