#!/bin/bash
for cfg in "4 3" "2 4" "2 6" "8 2" "4 4" "1 8"; do
  set -- $cfg
  GPSIG_FUSED=1 GPSIG_FUSED_R=$1 GPSIG_FUSED_S=$2 timeout 200 python bench.py --kernel linear --steps 3 --warmup 1 --no-cpu-baseline > gpurun_out/ft.json 2> gpurun_out/ft.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ft.json")); print("R=$1 S=$2 linear ms %.1f parity %.1e"%(d["ms_per_step"], d["parity"]["max_abs_err_over_max_abs_ref"]))
except Exception as e:
    print("R=$1 S=$2 FAILED", open("gpurun_out/ft.err").read()[-300:])
PY
done
