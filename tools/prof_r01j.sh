# round 1, tenth batch: whole GPU suite, smoke, difc / difp timing (fsum fused), default 1-GPU bench line,
# launch list of a small default-shaped run, ncu full of the difp solve.
set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/t_j.log 2>&1; echo "gpu suite rc=$?"; tail -5 gpurun_out/t_j.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_j.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/smoke_j.log
timeout 300 python tools/difc_sweep.py | tee gpurun_out/difc_sweep4.txt
timeout 1500 python bench.py > gpurun_out/bench_default_r01c.json 2> gpurun_out/bench_default_r01c.err; echo "bench rc=$?"; tail -c 400 gpurun_out/bench_default_r01c.err
A="python bench.py --cols 500 --steps 2 --warmup 1 --spinup 2 --no-cpu-baseline --no-e2e --kon-layers 2000 --bins-layers 2960"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r01c.csv $A > gpurun_out/ncu_launch_j.log 2>&1
echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:'difc_solve' -s 8 -c 3 -o gpurun_out/prof_r01j python tools/difc_sweep.py > gpurun_out/ncu_r01j.log 2>&1
echo "ncu rc=$?"
ncu -i gpurun_out/prof_r01j.ncu-rep --page raw --csv > gpurun_out/prof_r01j_raw.csv 2>/dev/null
rm -f gpurun_out/prof_r01j.ncu-rep
