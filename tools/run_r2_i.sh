#!/bin/bash
TAG=${1:-r2i}
timeout 120 python tools/debug_tc.py > gpurun_out/${TAG}_tc.log 2>&1; echo "debug_tc rc=$?"; tail -12 gpurun_out/${TAG}_tc.log
nvidia-smi --query-gpu=utilization.gpu,memory.used --format=csv,noheader
