#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
timeout 300 compute-sanitizer --tool racecheck --print-limit 400 python tools/racecheck_bf16.py > $O/ae_racecheck_bf16.log 2>&1; echo "racecheck rc=$?"
grep -E "RACECHECK SUMMARY|bf16 step done|Error" $O/ae_racecheck_bf16.log | sed 's/+0x[0-9a-f]*//' | sort | uniq -c | sort -rn | head -20
