#!/bin/bash
# A/B bench of library variants (build/variants/*.so) and env knobs: tools/gpu_ab.sh tag spec...
# spec = variant[:ENV=VAL[,ENV=VAL]]
tag=$1; shift
mkdir -p gpurun_out
for spec in "$@"; do
  v=${spec%%:*}; envs=""; [ "$spec" != "$v" ] && envs=$(echo "${spec#*:}" | tr ',' ' ')
  name=$(echo "$spec" | tr ':=,' '___')
  env $envs STGCN_B200_LIB=$PWD/build/variants/$v.so python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2> gpurun_out/${tag}_$name.err | grep '^{' > gpurun_out/${tag}_$name.json
  python - <<P
import json
try:
    d=json.load(open("gpurun_out/${tag}_$name.json"))
    print("$spec", round(d["value"]), "samples/s", round(d["ms_per_step"],4), "ms | e2e", round(d["e2e"]["value"]))
except Exception as e:
    print("$spec FAILED", e); print(open("gpurun_out/${tag}_$name.err").read()[-600:])
P
done
