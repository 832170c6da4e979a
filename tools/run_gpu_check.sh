#!/bin/bash
# usage: tools/run_gpu_check.sh TAG [ncu]   -- GPU tests, bench lines, optional ncu captures into gpurun_out/TAG_*
TAG=$1
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/${TAG}_gputests.log
for k in rbf linear; do timeout 600 python bench.py --kernel $k --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_cfg4_$k.json 2> gpurun_out/${TAG}_bench_cfg4_$k.err; done
timeout 300 python bench.py --workload cfg2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_cfg2.json 2> gpurun_out/${TAG}_bench_cfg2.err
if [ "$2" = "ncu" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sigkern_fo_ -s 20 -c 1 -f -o gpurun_out/${TAG}_prof_recursion python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1
fi
tail -3 gpurun_out/${TAG}_gputests.log
python - <<PY
import json
for f in ("cfg4_rbf","cfg4_linear","cfg2"):
    try:
        d=json.load(open("gpurun_out/${TAG}_bench_%s.json"%f))
        print(f, "value %.3e e2e %.3e ms %.2f roofline %.3f (%.0f GB/s) stages %s"%(d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["achieved"], {k:round(v["ms_per_step"],2) for k,v in d["stages"].items()}))
    except Exception as e:
        print(f, "FAILED", e); print(open("gpurun_out/${TAG}_bench_%s.err"%f).read()[-1500:])
PY
