#!/bin/bash
# Final round evidence: bench lines, reference arm, ncu launch list + full captures, other configs.  Output: gpurun_out/$TAG_*
TAG=$1
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/${TAG}_gputests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err
timeout 600 python bench.py --kernel linear --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_n1_linear.json 2>> gpurun_out/${TAG}_bench_n1.err
timeout 600 python bench.py --workload cfg2 --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_n1_cfg2.json 2>> gpurun_out/${TAG}_bench_n1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sigkern_fo_stream -s 20 -c 1 -f -o gpurun_out/${TAG}_prof_recursion python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/${TAG}_ncu_rec.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:delta_producer_fast -s 20 -c 1 -f -o gpurun_out/${TAG}_prof_producer_rbf python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/${TAG}_ncu_prod.log 2>&1
timeout 900 python tools/bench_configs.py --cfg all --steps 5 > gpurun_out/${TAG}_configs.jsonl 2> gpurun_out/${TAG}_configs.err
tail -2 gpurun_out/${TAG}_gputests.log; tail -2 gpurun_out/${TAG}_smoke.log
python - <<PY
import json
for f in ("n1","n1_linear","n1_cfg2","reference"):
    try:
        d=json.loads(open("gpurun_out/${TAG}_bench_%s.json"%f).read().strip().splitlines()[-1])
        print(f, "value %.3e e2e %.3e ms %.2f"%(d["value"], d["e2e"]["value"], d["ms_per_step"]), "roofline", d.get("roofline",{}).get("frac"), "cpu", (d.get("cpu_baseline") or {}).get("value"), "launches", d.get("gpu_launches"), "clocks", d.get("clocks"))
    except Exception as e:
        print(f, "FAILED", e)
PY
tail -3 gpurun_out/${TAG}_bench_n1.err
