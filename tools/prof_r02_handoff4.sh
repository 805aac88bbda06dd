cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_zz_onchip.py tests/test_gpu_rates.py tests/test_gpu_multi.py tests/test_gpu_b1.py -m gpu -q -x ) > gpurun_out/r02_handoff_host_tests.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/r02_handoff_host_tests.log
timeout 600 python bench.py --no-extras --no-cpu-baseline --no-bins > gpurun_out/r02_handoff_host_bench.json 2> gpurun_out/r02_handoff_host_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_handoff_host_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r02_handoff_host_bench.json') if l.startswith('{')][-1])
print('value %.4g e2e %.4g rconst_path %.4g'%(d['value'], d['e2e']['value'], d['e2e'].get('rconst_path',{}).get('value',0)), 'ms', d['ms_per_step'], d['e2e']['ms_per_step'], 'parity', d['parity']['ok'])
PY
