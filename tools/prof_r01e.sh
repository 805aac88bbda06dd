# round 1, fifth batch: difc (merged launch, warp-per-column coefficients) and difp: GPU tests, occupancy sweep, ncu.
set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_difc.py -x -q > gpurun_out/t_difc.log 2>&1; echo "difc tests rc=$?"; tail -15 gpurun_out/t_difc.log
for r in 16 8 6 4 3 2; do MISTRA_DIFC_CTAS_PER_SM=$r timeout 300 python tools/difc_sweep.py; done 2>&1 | grep CTAs | tee gpurun_out/difc_sweep.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'difc_|difp_' -c 8 -o gpurun_out/prof_r01e python tools/difc_sweep.py > gpurun_out/ncu_r01e.log 2>&1
echo "ncu rc=$?"
ncu -i gpurun_out/prof_r01e.ncu-rep --page raw --csv > gpurun_out/prof_r01e_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_r01e.ncu-rep --page source --csv > gpurun_out/prof_r01e_src.csv 2>/dev/null
rm -f gpurun_out/prof_r01e.ncu-rep
