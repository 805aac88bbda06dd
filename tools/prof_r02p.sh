cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python tools/oc_stress.py 40 > gpurun_out/r02p_stress.txt 2>&1; tail -40 gpurun_out/r02p_stress.txt
