#!/bin/bash
# round 2, call Q: the whole acceptance pass on one GPU -- tests, smoke, bench lines of every workload (default line with
# cpu_baseline + pipeline), reference arm, ncu launch list of the default bench, ncu --set full of the two hot kernels
TAG=${1:-r2q}
timeout 1800 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/${TAG}_gputests_full.log 2>&1
grep -E "AssertionError: |Error|passed|failed" gpurun_out/${TAG}_gputests_full.log | sort | uniq -c | sort -rn | head -12
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
show() {
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_$1.json").read().strip().splitlines()[-1])
    print("$1", "value %.3e e2e %.3e ms %.3f"%(d["value"], d["e2e"]["value"], d["ms_per_step"]), "fused", round(d.get("stages",{}).get("fused",{}).get("ms_per_step",0),2), "tens", round(d.get("stages",{}).get("tens",{}).get("ms_per_step",0),3), "roof", d["roofline"]["bound"] if d.get("roofline") else None, round(d["roofline"]["frac"],3) if d.get("roofline") else None, "cpu", (d.get("cpu_baseline") or {}).get("value"), "pipeline", (d.get("pipeline") or {}).get("ms_per_step"))
except Exception as e:
    print("$1 FAILED", e); print(open("gpurun_out/${TAG}_$1.err").read()[-1500:])
PY
}
timeout 900 python bench.py > gpurun_out/${TAG}_bench_default.json 2> gpurun_out/${TAG}_bench_default.err; show bench_default
for wl in cfg1 cfg2 cfg3 cfg5; do
  timeout 900 python bench.py --workload $wl > gpurun_out/${TAG}_bench_$wl.json 2> gpurun_out/${TAG}_bench_$wl.err; show bench_$wl
done
timeout 900 python bench.py --kernel linear > gpurun_out/${TAG}_bench_cfg4_linear.json 2> gpurun_out/${TAG}_bench_cfg4_linear.err; show bench_cfg4_linear
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; tail -c 600 gpurun_out/${TAG}_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-pipeline > gpurun_out/${TAG}_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sigkern_warpfused -s 2 -c 1 -f -o gpurun_out/${TAG}_prof_wf_rbf python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-pipeline > gpurun_out/${TAG}_ncu_wf_rbf.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tens_seq_tc -s 2 -c 1 -f -o gpurun_out/${TAG}_prof_tc python bench.py --workload cfg3 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_tc.log 2>&1
ls -la gpurun_out/ | tail -5
