#!/bin/bash
# One GPU-box pass: parity tests, bench line, ncu launch list, ncu --set full of the fused graph-conv kernels.
# usage: tools/gpu_check.sh <tag>      (outputs under gpurun_out/<tag>_*)
tag=${1:-run}
mkdir -p gpurun_out
(time timeout 420 python -m pytest tests -m gpu -x -q) > gpurun_out/${tag}_pytest.log 2>&1
tail -3 gpurun_out/${tag}_pytest.log
(time python bench.py --steps 20 --warmup 5) > gpurun_out/${tag}_bench.log 2>&1
grep '^{' gpurun_out/${tag}_bench.log > gpurun_out/${tag}_bench.json
cut -c1-400 gpurun_out/${tag}_bench.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 220 --csv \
  --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-profile \
  > gpurun_out/${tag}_ncu1.log 2>&1
if [ -n "$2" ]; then
timeout 240 ncu --set full --clock-control none --import-source on -k "regex:$2" -s ${3:-4} -c ${4:-4} -f -o gpurun_out/${tag}_full \
  python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-profile > gpurun_out/${tag}_ncu2.log 2>&1
fi
ls -la gpurun_out | tail -12
