cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_liq_reference.py tests/test_rconst_reference.py tests/test_gpu_rconst.py -m gpu -q > gpurun_out/r02_liq_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r02_liq_tests.log
timeout 300 python - <<'PY' > gpurun_out/r02_liq_bench.txt 2>&1
import numpy as np, torch, sys
sys.path.insert(0, ".")
from mistra_b200 import liq
n = 98000
r = np.random.default_rng(1)
t = torch.from_numpy(r.uniform(240, 300, n)).cuda()
cw = torch.from_numpy(10.0 ** r.uniform(-12, -7, (n, 2))).cuda()
conv2 = 1.0 / (1000.0 * cw)
xg = torch.from_numpy(r.uniform(0.3, 1.5, (n, 2, 55))).cuda()
for it in range(4):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = liq.tables_device(1, t, conv2.contiguous(), xg); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
byt = n * (262 * 3 + 2 * 262 * 2) * 8 * 2 + n * (1 + 2 + 110) * 8       # memset + kernel writes, inputs
print("liq tables (aer): %d layers %.3f ms = %.1f M layers/s, %.0f GB/s of output + input bytes" % (n, ms, n / ms / 1e3, byt / ms / 1e6))
PY
cat gpurun_out/r02_liq_bench.txt
