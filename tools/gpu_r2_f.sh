#!/bin/bash
# round 2, GPU call F: validate residual-MMA restriction + deeper fb0 pipelines; tap timeline probe.
mkdir -p gpurun_out; O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 120 > $O/f_tests.log 2>&1; echo "tests rc=$?" | tee -a $O/f_tests.log
tail -8 $O/f_tests.log
b() { name=$1; shift; timeout 200 python bench.py --no-extras --no-cpu-baseline --steps 20 --warmup 5 "$@" > $O/f_bench_$name.json 2> $O/f_bench_$name.err; echo "bench $name rc=$?"; tail -2 $O/f_bench_$name.err; }
b base
STGCN_B200_LIB=$PWD/build/variants/tl.so timeout 200 python tools/tap_timeline_probe.py > $O/f_timeline.txt 2>&1; echo "timeline rc=$?"; cat $O/f_timeline.txt
for f in $O/f_bench*.json; do echo $f; python -c "
import json,sys
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline_step']['frac'])
    for k in d['top_kernels'][:45]: print('   ', round(k['ms_per_step']*1000,1), k['key'][:90])
except Exception as e: print('ERR', e)
"; done
