#!/bin/bash
TAG=${1:-r2c}
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/${TAG}_gputests.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sigkern_warpfused -s 2 -c 1 -f -o gpurun_out/${TAG}_prof_wf_rbf python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${TAG}_ncu_wf_rbf.log 2>&1
cat gpurun_out/${TAG}_gputests.log
