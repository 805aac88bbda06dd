cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 600 python bench.py --no-extras --no-e2e --mechs gas --cols 10000 ) > gpurun_out/r05_bench_nextrows.json 2> gpurun_out/r05_bench_nextrows.err; echo "rc=$?"; tail -3 gpurun_out/r05_bench_nextrows.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r05_bench_nextrows.json') if l.startswith('{')][-1])
for k,x in d['next_rows'].items(): print(k,'%.4g'%x['value'],x['unit'],'ms %.3f'%x['ms_per_step'],'frac %.3f'%x['roofline']['frac'], x.get('cpu_baseline',{}).get('value'))
PY
