#!/bin/bash
TAG=${1:-r2m}
timeout 600 python -m pytest tests/test_gpu_algs.py tests/test_gpu_parallel.py tests/test_gpu_conditioning.py -m gpu -q 2>&1 | tail -6
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tens_seq_tc -s 2 -c 1 -f -o gpurun_out/${TAG}_prof_tc python bench.py --workload cfg3 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_tc.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_tc.log | cut -c1-200
python - <<'PY'
import time, numpy as np, torch, sys
sys.path.insert(0, '.')
import bench
from gpsig_b200 import kernels
# higher-order path with a number: configs[1] shape with order = M (notebook semantics)
N, L, d, M = 1024, 64, 6, 4
X = torch.as_tensor(bench.synth_X(N, L, d).astype(np.float32), device="cuda")
k = kernels.SignatureLinear(L * d, d, M, order=M, normalization=False)
for _ in range(2): K = k.K(X)
torch.cuda.synchronize(); t0 = time.time()
for _ in range(3): K = k.K(X)
torch.cuda.synchronize(); dt = (time.time() - t0) / 3
print("higher-order K(X,X) N=%d L=%d d=%d M=%d order=%d: %.1f ms/step = %.3e pairs/s" % (N, L, d, M, M, dt * 1e3, N * N / dt))
PY
