#!/bin/bash
# round 2, GPU call AA: ln_bwd_sums_pg_kernel with cp.async double buffering
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 300 > $O/aa_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/aa_tests.log | cut -c1-300
timeout 200 python bench.py --no-extras --no-cpu-baseline --steps 20 --warmup 5 > $O/aa_bench_base.json 2> $O/aa_bench_base.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('$O/aa_bench_base.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'])
for k in d['top_kernels'][:24]: print('   ', round(k['ms_per_step']*1000,1), k['key'][:90])
"
