#!/bin/bash
# round 2: N-rank runs (gpurun --gpus N -- './tools/run_r2_multi.sh N TAG'): parallel tests, bench lines of cfg4 / cfg3 / cfg5
N=$1; TAG=${2:-r2multi}
[ -n "$SKIPTESTS" ] || timeout 600 python -m pytest tests/test_gpu_parallel.py -m gpu -q 2>&1 | tail -3
PORT=29500
for wl in ${WLS:-cfg4 cfg3 cfg5}; do
  PORT=$((PORT+1))
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $N --steps 10 --warmup 3 --workload $wl --no-pipeline > gpurun_out/${TAG}_${wl}_n${N}.json 2> gpurun_out/${TAG}_${wl}_n${N}.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_${wl}_n${N}.json").read().strip().splitlines()[-1])
    print("$wl N=$N", "value %.3e e2e %.3e ms %.3f e2e_ms %.3f"%(d["value"], d["e2e"]["value"], d["ms_per_step"], d["e2e"]["ms_per_step"]), "launches", d["gpu_launches"], "per_rank_kernel_ms", d.get("per_rank_kernel_ms"), "parity", {k:(v if not isinstance(v,dict) else v.get("max_abs_err_over_max_abs_ref")) for k,v in d["parity"].items()}, "clocks", d["clocks"])
except Exception as e:
    print("$wl N=$N FAILED", e); print(open("gpurun_out/${TAG}_${wl}_n${N}.err").read()[-2500:])
PY
done
