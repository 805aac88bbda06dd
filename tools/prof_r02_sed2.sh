cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sed.py -m gpu -q -x > gpurun_out/r02_sed_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r02_sed_tests.log
timeout 300 python tools/sed_bench.py 256 > gpurun_out/r02_sed_bench.txt 2>&1; cat gpurun_out/r02_sed_bench.txt
timeout 300 python tools/sed_bench.py 1024 2>&1 | tee -a gpurun_out/r02_sed_bench.txt
