cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for s in 0 1; do
MISTRA_KPP_SPLIT=$s timeout 900 python bench.py --cols 6000 --mechs aer --no-e2e --no-extras --no-bins --no-cpu-baseline > gpurun_out/r02z_split$s.json 2> gpurun_out/r02z_split$s.err; python -c "
import json; l=json.loads(open('gpurun_out/r02z_split$s.json').read().strip().splitlines()[-1]); print('split=$s value', l['value'], 'ms', l['ms_per_step'])"
done
timeout 600 python tools/oc_bench.py aer 40 1 0 > gpurun_out/r02z_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ros3_onchip_a -s 2 -c 1 -o gpurun_out/r02z_onchip_aer python tools/oc_bench.py aer 40 1 0 > gpurun_out/r02z_ncu.log 2>&1
tail -2 gpurun_out/r02z_ncu.log
timeout 600 python bench.py --cols 300 --steps 2 --warmup 1 --tot-cells 9800 --no-cpu-baseline --kon-layers 300 --bins-layers 1480 > gpurun_out/r02z_small.json 2>/dev/null &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02z_launches.csv python bench.py --cols 300 --steps 2 --warmup 1 --tot-cells 9800 --no-cpu-baseline --kon-layers 300 --bins-layers 1480 > gpurun_out/r02z_ncu_small.log 2>&1
wc -l gpurun_out/r02z_launches.csv
