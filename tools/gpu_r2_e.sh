#!/bin/bash
# round 2, GPU call E: tap kernel with bias/residual on the tensor pipe, device GSO, syn2048 workload.
mkdir -p gpurun_out; O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 120 > $O/e_tests.log 2>&1; echo "tests rc=$?" | tee -a $O/e_tests.log
tail -12 $O/e_tests.log
b() { name=$1; shift; timeout 200 python bench.py --no-extras --no-cpu-baseline --steps 20 --warmup 5 "$@" > $O/e_bench_$name.json 2> $O/e_bench_$name.err; echo "bench $name rc=$?"; tail -2 $O/e_bench_$name.err; }
b base
b metrla --workload metrla
b pemsbay --workload pemsbay --batch 256
b syn2048 --workload syn2048 --batch 32 --steps 5 --warmup 3
b x3 --precision tf32x3
for f in $O/e_bench*.json; do echo $f; python -c "
import json,sys
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline_step']['frac'])
    for k in d['top_kernels'][:10]: print('   ', round(k['ms_per_step']*1000,1), k['key'][:90])
except Exception as e: print('ERR', e)
"; done
