cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
for h in 0 12; do
  echo "== 2 GPUs weak, MISTRA_KPP_HANDOFF=$h"
  MISTRA_KPP_HANDOFF=$h timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-extras --no-cpu-baseline --no-bins --no-e2e > gpurun_out/r02_handoff_bench2_$h.json 2> gpurun_out/r02_handoff_bench2_$h.err
  python - $h <<'PY'
import json,sys
d=json.loads([l for l in open('gpurun_out/r02_handoff_bench2_%s.json'%sys.argv[1]) if l.startswith('{')][-1])
print('value %.4g ms %.1f'%(d['value'], d['ms_per_step']), d.get('per_mechanism'), d['diagnostics'], d.get('parity',{}).get('ok'), 'launches', d.get('gpu_launches'))
PY
done
