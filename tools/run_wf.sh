#!/bin/bash
timeout 400 python -m pytest tests/test_gpu_large.py tests/test_gpu_kernels.py -m gpu -q -x 2>&1 | tail -2
for v in 0 1; do
GPSIG_WARPFUSED_RBF12=$v timeout 200 python bench.py --kernel rbf --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('rbf12=$v ms %.1f value %.3e parity %.1e clocks %s'%(d['ms_per_step'], d['value'], d['parity']['max_abs_err_over_max_abs_ref'], d['clocks']['sm_mhz']))"
done
for w in 12 10 8; do
GPSIG_WARPFUSED_WARPS=$w timeout 200 python bench.py --kernel linear --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('linear warps=$w ms %.1f value %.3e'%(d['ms_per_step'], d['value']))"
done
