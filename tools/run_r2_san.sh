#!/bin/bash
# round 2: compute-sanitizer memcheck over small invocations of every tuned kernel (tools/sanitize_small.py)
TAG=${1:-r2san}
timeout 120 python tools/sanitize_small.py 2>&1 | tail -3
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 1 python tools/sanitize_small.py > gpurun_out/${TAG}_memcheck.log 2>&1; echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|Invalid|sanitize_small ok|Error" gpurun_out/${TAG}_memcheck.log | head -20
timeout 1500 compute-sanitizer --tool initcheck --error-exitcode 1 python tools/sanitize_small.py > gpurun_out/${TAG}_initcheck.log 2>&1; echo "initcheck rc=$?"
grep -E "ERROR SUMMARY|Uninitialized|sanitize_small ok" gpurun_out/${TAG}_initcheck.log | sort | uniq -c | head -10
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report analysis python tools/sanitize_small.py > gpurun_out/${TAG}_racecheck.log 2>&1; echo "racecheck rc=$?"
grep -E "RACECHECK SUMMARY|hazard|sanitize_small ok" gpurun_out/${TAG}_racecheck.log | cut -c1-220 | sort | uniq -c | sort -rn | head -12
