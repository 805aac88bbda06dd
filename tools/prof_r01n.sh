set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/t_n.log 2>&1; echo "gpu suite rc=$?"; tail -6 gpurun_out/t_n.log
