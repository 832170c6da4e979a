#!/bin/bash
TAG=${1:-r2d}
timeout 900 python -m pytest tests/test_gpu_conditioning.py -m gpu -q 2>&1 | grep -E "AssertionError: |passed|failed" > gpurun_out/${TAG}_cond.log
cat gpurun_out/${TAG}_cond.log
