cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python tools/oc_diag.py > gpurun_out/r02n_diag.txt 2>&1; cat gpurun_out/r02n_diag.txt
