#!/bin/bash
# compute-sanitizer memcheck of smoke() on the final tree (KV ordering on the producer warp: new shared words, done words)
T=${1:-r2san2}
mkdir -p gpurun_out
timeout 200 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/${T}_memcheck.log
grep -E "smoke ok|ERROR SUMMARY|rc=|Invalid|Error" gpurun_out/${T}_memcheck.log | head -12
