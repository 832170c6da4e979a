#!/bin/bash
TAG=${1:-r2l}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tens_seq_tc -s 2 -c 1 -f -o gpurun_out/${TAG}_prof_tc python bench.py --workload cfg3 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_tc.log 2>&1
tail -3 gpurun_out/${TAG}_ncu_tc.log | cut -c1-300
