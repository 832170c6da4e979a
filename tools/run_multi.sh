#!/bin/bash
# usage: tools/run_multi.sh TAG N   -- multi-GPU parity test + bench at N ranks (run under gpurun --gpus N)
TAG=$1; N=$2
timeout 600 python -m pytest tests/test_gpu_parallel.py -m gpu -q -x 2>&1 | tail -5 > gpurun_out/${TAG}_partest_${N}.log
PORT=29517
for k in rbf linear; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus $N --kernel $k --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_${k}_n${N}.json 2> gpurun_out/${TAG}_bench_${k}_n${N}.err
PORT=$((PORT+1))
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $PORT bench.py --impl reference --gpus $N --steps 1 --warmup 0 > gpurun_out/${TAG}_bench_ref_n${N}.json 2> gpurun_out/${TAG}_bench_ref_n${N}.err
tail -2 gpurun_out/${TAG}_partest_${N}.log
python - <<PY
import json
for k in ("rbf","linear","ref"):
    f="gpurun_out/${TAG}_bench_%s_n${N}.json"%k
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(k, "n_gpus", d["n_gpus"], "value %.3e e2e %.3e ms %.2f"%(d["value"], d["e2e"]["value"], d["ms_per_step"]), d.get("clocks"), d.get("stages") and {a:round(b["ms_per_step"],2) for a,b in d["stages"].items()})
    except Exception as e:
        print(k, "FAILED", e); print(open(f.replace(".json",".err")).read()[-2000:])
PY
