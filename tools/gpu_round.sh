#!/bin/bash
# One GPU-box pass for a batch of changes: parity tests (bisecting over the A/B knobs when they fail), then A/B benches.
# OPTIN="KNOB=1 ..." (env): extra parity passes with opt-in features switched on
# usage: tools/gpu_round.sh tag "knob1 knob2 ..." spec...     (specs as in tools/gpu_ab.sh)
tag=$1; knobs=$2; shift 2
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; rc=$?
tail -4 gpurun_out/${tag}_pytest.log
if [ $rc -ne 0 ]; then
  grep -E "^(FAILED|ERROR)|Error|error" gpurun_out/${tag}_pytest.log | head -8 | cut -c1-300
  for k in $knobs; do
    env $k timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_$k.log 2>&1
    echo "with $k: rc=$? $(tail -1 gpurun_out/${tag}_pytest_$k.log)"
  done
fi
for k in $OPTIN; do
  env $k timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_$k.log 2>&1
  echo "opt-in $k: rc=$? $(tail -1 gpurun_out/${tag}_pytest_$k.log)"
done
tools/gpu_ab.sh $tag "$@"
