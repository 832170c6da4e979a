#!/bin/bash
# round 2, call A: FP32 pipe microbenchmark + ncu --set full of the RBF warp-fused instantiation at cfg4 (baseline)
./tools/ubench/pipes > gpurun_out/r2a_pipes.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sigkern_warpfused -s 1 -c 1 -f -o gpurun_out/r2a_prof_wf_rbf python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2a_ncu_wf_rbf.log 2>&1
tail -5 gpurun_out/r2a_ncu_wf_rbf.log
cat gpurun_out/r2a_pipes.log
