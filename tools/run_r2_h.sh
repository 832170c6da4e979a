#!/bin/bash
TAG=${1:-r2h}
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_gputests_full.log 2>&1
grep -E "AssertionError: |Error|passed|failed" gpurun_out/${TAG}_gputests_full.log | sort | uniq -c | sort -rn | head -30
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench_n1.err
python - <<PY
import json
for f in ("n1","reference"):
    try:
        d=json.loads(open("gpurun_out/${TAG}_bench_%s.json"%f).read().strip().splitlines()[-1])
        print(f, "value %.3e e2e %.3e ms %.2f"%(d["value"], d["e2e"]["value"], d["ms_per_step"]), "roofline", {k:v for k,v in (d.get("roofline") or {}).items() if k in ("bound","frac","achieved","peak")}, "pipeline", (d.get("pipeline") or {}).get("ms_per_step"), (d.get("pipeline") or {}).get("roofline",{}).get("frac"), "cpu", (d.get("cpu_baseline") or {}).get("value"), "launches", d.get("gpu_launches"), "clocks", d.get("clocks"), "parity", d.get("parity"))
    except Exception as e:
        print(f, "FAILED", e)
PY
tail -3 gpurun_out/${TAG}_bench_n1.err
