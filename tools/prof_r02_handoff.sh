cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests/test_gpu_zz_onchip.py tests/test_gpu_parity.py tests/test_gpu_multi.py tests/test_gpu_rates.py -m gpu -q -x ) > gpurun_out/r02_handoff_tests.log 2>&1; echo "tests rc=$?"; tail -25 gpurun_out/r02_handoff_tests.log
for h in 0 12; do
  echo "== MISTRA_KPP_HANDOFF=$h"
  MISTRA_KPP_HANDOFF=$h timeout 900 python bench.py --no-extras --no-cpu-baseline --no-bins --no-e2e > gpurun_out/r02_handoff_bench_$h.json 2> gpurun_out/r02_handoff_bench_$h.err
  python - $h <<'PY'
import json,sys
d=json.loads([l for l in open('gpurun_out/r02_handoff_bench_%s.json'%sys.argv[1]) if l.startswith('{')][-1])
print('value %.4g ms %.1f'%(d['value'], d['ms_per_step']), d.get('per_mechanism'), d['diagnostics'], d.get('parity',{}).get('ok'), 'launches', d.get('gpu_launches'))
PY
done
