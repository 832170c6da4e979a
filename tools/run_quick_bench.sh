#!/bin/bash
# usage: tools/run_quick_bench.sh TAG  -- bench lines only (no tests)
TAG=$1
for k in rbf linear; do timeout 600 python bench.py --kernel $k --steps 4 --warmup 2 --no-cpu-baseline > gpurun_out/${TAG}_bench_cfg4_$k.json 2> gpurun_out/${TAG}_bench_cfg4_$k.err; done
timeout 300 python bench.py --workload cfg2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_cfg2.json 2> gpurun_out/${TAG}_bench_cfg2.err
python - <<PY
import json
for f in ("cfg4_rbf","cfg4_linear","cfg2"):
    try:
        d=json.load(open("gpurun_out/${TAG}_bench_%s.json"%f))
        print(f, "value %.3e e2e %.3e ms %.2f parity %.1e clocks %s stages %s"%(d["value"], d["e2e"]["value"], d["ms_per_step"], d["parity"]["max_abs_err_over_max_abs_ref"], d["clocks"]["sm_mhz"], {k:round(v["ms_per_step"],2) for k,v in d["stages"].items() if v["ms_per_step"]>0.3}))
    except Exception as e:
        print(f, "FAILED", e); print(open("gpurun_out/${TAG}_bench_%s.err"%f).read()[-1500:])
PY
