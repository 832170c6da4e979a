#!/bin/bash
TAG=$1
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/${TAG}_gputests.log
tail -3 gpurun_out/${TAG}_gputests.log
timeout 900 python tools/bench_configs.py --cfg all --steps 5 > gpurun_out/${TAG}_configs.jsonl 2> gpurun_out/${TAG}_configs.err
cat gpurun_out/${TAG}_configs.jsonl; tail -5 gpurun_out/${TAG}_configs.err
