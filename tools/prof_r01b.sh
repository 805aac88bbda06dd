# ncu captures of round 1 (second batch): launch list of a small default-shaped run, full sets of the
# particle-grid / next-row kernels and of ros3_kernel_a.  The .ncu-rep files exceed gpurun's 64 MiB
# return limit, so only the raw-page CSV comes back.  Run under gpurun: bash tools/prof_r01b.sh
set -x
cd $GRAFT_REPO_ROOT
A="python bench.py --cols 500 --steps 2 --warmup 1 --spinup 2 --no-cpu-baseline --no-e2e --kon-layers 2000 --bins-layers 2960"
$A > gpurun_out/plain_small.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01b.csv $A > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"
$A > gpurun_out/plain_small2.log 2>&1 && ncu --set full --clock-control none -k regex:'kon_subkon|konc_kernel|cwrc_kernel|rconst_kernel' -c 12 -o gpurun_out/prof_small $A > gpurun_out/ncu_small.log 2>&1
echo "small full rc=$?"
ncu -i gpurun_out/prof_small.ncu-rep --page raw --csv > gpurun_out/prof_small_raw.csv 2>/dev/null
rm -f gpurun_out/prof_small.ncu-rep
B="python bench.py --mechs aer --cols 1000 --steps 1 --warmup 1 --spinup 3 --no-cpu-baseline --no-e2e --no-bins"
$B > gpurun_out/plain_aer.log 2>&1 && ncu --set full --clock-control none -k regex:ros3_kernel -s 4 -c 1 -o gpurun_out/prof_aer2 $B > gpurun_out/ncu_aer2.log 2>&1
echo "aer full rc=$?"
ncu -i gpurun_out/prof_aer2.ncu-rep --page raw --csv > gpurun_out/prof_aer2_raw.csv 2>/dev/null
rm -f gpurun_out/prof_aer2.ncu-rep
ls -la gpurun_out/
