#!/bin/bash
# round 2, GPU call V: umma_fb2 -- per-input-step data-gradient accumulators (one reader), deterministic group sums,
# LayerNorm parameter gradients on the helper stream
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 120 > $O/v_tests.log 2>&1; echo "tests rc=$?"; tail -4 $O/v_tests.log | cut -c1-300
b() { name=$1; shift; timeout 200 python bench.py --no-extras --no-cpu-baseline --steps 20 --warmup 5 "$@" > $O/v_bench_$name.json 2> $O/v_bench_$name.err; echo "bench $name rc=$?"; tail -2 $O/v_bench_$name.err; 
python -c "
import json
d=json.loads(open('$O/v_bench_$name.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'])
for k in d['top_kernels'][:32]: print('   ', round(k['ms_per_step']*1000,1), k['key'][:90])
"; }
b base
