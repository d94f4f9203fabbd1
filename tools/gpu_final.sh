#!/bin/bash
# Final measurement pass of a round on one B200: parity tests, the bench line (with the CPU baseline), the reference arm,
# the other single-GPU workloads, an ncu launch list and one `ncu --set full` capture of the dominant kernels.
# usage: tools/gpu_final.sh <tag>     (outputs under gpurun_out/<tag>_*)
tag=${1:-final}
mkdir -p gpurun_out
(time timeout 300 python -m pytest tests -m gpu -x -q) > gpurun_out/${tag}_pytest.log 2>&1; tail -4 gpurun_out/${tag}_pytest.log
python bench.py --steps 20 --warmup 5 2> gpurun_out/${tag}_bench.err | grep '^{' > gpurun_out/${tag}_bench.json; cut -c1-330 gpurun_out/${tag}_bench.json
python bench.py --impl reference --steps 5 --warmup 3 2>/dev/null | grep '^{' > gpurun_out/${tag}_ref.json; cut -c1-200 gpurun_out/${tag}_ref.json
for w in metrla pemsbay; do
  python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | grep '^{' > gpurun_out/${tag}_bench_$w.json; cut -c1-200 gpurun_out/${tag}_bench_$w.json
done
python bench.py --precision fp32 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | grep '^{' > gpurun_out/${tag}_bench_fp32.json; cut -c1-200 gpurun_out/${tag}_bench_fp32.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 240 --csv \
  --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-profile \
  > gpurun_out/${tag}_ncu1.log 2>&1
timeout 280 ncu --set full --clock-control none --import-source on \
  -k "regex:umma_tap_kernel|ln_gate_bwd_pipe|umma_cheb|smallc1_gate_wgrad|umma_wgrad_kernel" -s 30 -c 12 -f -o gpurun_out/${tag}_full \
  python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline --no-profile > gpurun_out/${tag}_ncu2.log 2>&1
ls -la gpurun_out | grep ${tag}_ | awk '{print $5, $9}'
