#!/bin/bash
TAG=$1
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/${TAG}_gputests.log
tail -3 gpurun_out/${TAG}_gputests.log
for cfg in "8 4 3" "10 4 2" "11 4 2" "11 2 5" "10 3 3" "11 3 3" "8 2 6" "9 4 3"; do
  set -- $cfg
  for k in linear; do
  GPSIG_STREAM_NCW=$1 GPSIG_STREAM_R=$2 GPSIG_STREAM_S=$3 timeout 300 python bench.py --kernel $k --steps 4 --warmup 2 --no-cpu-baseline > gpurun_out/${TAG}_tune.json 2> gpurun_out/${TAG}_tune.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_tune.json"))
    print("ncw=$1 R=$2 S=$3 $k: value %.3e ms %.1f recursion frac %.3f (%.0f GB/s) clocks %s stages %s"%(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["achieved"], d["clocks"]["sm_mhz"], {a:round(b["ms_per_step"],1) for a,b in d["stages"].items() if b["ms_per_step"]>0.5}))
except Exception as e:
    print("ncw=$1 R=$2 S=$3 FAILED", e, open("gpurun_out/${TAG}_tune.err").read()[-600:])
PY
  done
done
timeout 900 python tools/bench_configs.py --cfg all --steps 5 > gpurun_out/${TAG}_configs.jsonl 2> gpurun_out/${TAG}_configs.err
python - <<PY
import json
for ln in open("gpurun_out/${TAG}_configs.jsonl"):
    d=json.loads(ln); print(d["config"], "ms %.2f"%d["ms"], {k:round(v["ms_per_call"],2) for k,v in d["stages"].items()}, {k:v for k,v in d.items() if "relerr" in k})
PY
tail -4 gpurun_out/${TAG}_configs.err
