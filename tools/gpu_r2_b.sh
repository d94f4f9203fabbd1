#!/bin/bash
# round 2, GPU call B: re-validated tf32x3 kernels (two producer groups), micro-streams, clock sampler, N2/N3 GPU tests.
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 180 -x -k "x3 or tf32x3 or train or micro or windows" > $O/b_tests.log 2>&1; echo "tests rc=$?" | tee -a $O/b_tests.log
tail -15 $O/b_tests.log
b() { name=$1; shift; timeout 240 python bench.py --no-extras --no-cpu-baseline --steps 20 --warmup 5 "$@" > $O/b_bench_$name.json 2> $O/b_bench_$name.err; echo "bench $name rc=$?"; tail -2 $O/b_bench_$name.err; }
b base
b micro2 --micro-streams 2
b micro4 --micro-streams 4
b x3 --precision tf32x3
timeout 400 python bench.py > $O/b_bench.json 2> $O/b_bench.err; echo "bench rc=$?"; tail -3 $O/b_bench.err
for f in $O/b_bench*.json; do echo $f; python -c "
import json,sys
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])
    for k in d['top_kernels'][:6]: print('   ', round(k['ms_per_step']*1000,1), k['key'][:90])
except Exception as e: print('ERR', e)
"; done
