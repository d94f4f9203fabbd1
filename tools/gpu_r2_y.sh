#!/bin/bash
# round 2, GPU call Y: umma_fb2 with four dedicated E2 warps and the trimmed E1 arithmetic
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 300 > $O/y_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/y_tests.log | cut -c1-300
STGCN_B200_LIB=$PWD/build/variants/tl.so timeout 120 python tools/fb2_cycles_probe.py > $O/y_fb2_cycles.txt 2>&1; cat $O/y_fb2_cycles.txt
timeout 200 python bench.py --no-extras --no-cpu-baseline --steps 20 --warmup 5 > $O/y_bench_base.json 2> $O/y_bench_base.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('$O/y_bench_base.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'])
for k in d['top_kernels'][:16]: print('   ', round(k['ms_per_step']*1000,1), k['key'][:90])
"
