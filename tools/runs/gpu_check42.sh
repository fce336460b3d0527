#!/bin/bash
# compute-sanitizer memcheck over the kernels added in the second half of round 2
T=${1:-r2san}
mkdir -p gpurun_out
timeout 300 python tools/sanitize_new_kernels.py > gpurun_out/${T}_plain.log 2>&1; echo "plain rc=$?" >> gpurun_out/${T}_plain.log; tail -6 gpurun_out/${T}_plain.log
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 1 python tools/sanitize_new_kernels.py > gpurun_out/${T}_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/${T}_memcheck.log
grep -E "ok$|ERROR SUMMARY|rc=|Invalid|Error" gpurun_out/${T}_memcheck.log | head -20
