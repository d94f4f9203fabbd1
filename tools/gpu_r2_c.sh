#!/bin/bash
# round 2, GPU call C: fused first-block backward (umma_fb0), x3 kernel with 32-bit index math, training-side tests.
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 180 > $O/c_tests.log 2>&1; echo "tests rc=$?" | tee -a $O/c_tests.log
tail -25 $O/c_tests.log
b() { name=$1; shift; timeout 240 python bench.py --no-extras --no-cpu-baseline --steps 20 --warmup 5 "$@" > $O/c_bench_$name.json 2> $O/c_bench_$name.err; echo "bench $name rc=$?"; tail -2 $O/c_bench_$name.err; }
b base
b x3 --precision tf32x3
b metrla --workload metrla
b pemsbay --workload pemsbay --batch 256
for f in $O/c_bench*.json; do echo $f; python -c "
import json,sys
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])
    for k in d['top_kernels'][:8]: print('   ', round(k['ms_per_step']*1000,1), k['key'][:90])
except Exception as e: print('ERR', e)
"; done
