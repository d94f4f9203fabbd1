#!/bin/bash
# round 2, GPU call G: tap kernel with a TMA-store warp (no CTA-wide barriers in the epilogue), fb0 operand prefetch.
mkdir -p gpurun_out; O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 120 > $O/g_tests.log 2>&1; echo "tests rc=$?" | tee -a $O/g_tests.log
tail -8 $O/g_tests.log
b() { name=$1; shift; timeout 200 python bench.py --no-extras --no-cpu-baseline --steps 20 --warmup 5 "$@" > $O/g_bench_$name.json 2> $O/g_bench_$name.err; echo "bench $name rc=$?"; tail -2 $O/g_bench_$name.err; }
b base
b metrla --workload metrla
STGCN_B200_LIB=$PWD/build/variants/tl.so timeout 200 python tools/tap_timeline_probe.py > $O/g_timeline.txt 2>&1; echo "timeline rc=$?"; head -8 $O/g_timeline.txt | cut -c1-400
for f in $O/g_bench*.json; do echo $f; python -c "
import json,sys
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline_step']['frac'])
    for k in d['top_kernels'][:32]: print('   ', round(k['ms_per_step']*1000,1), k['key'][:90])
except Exception as e: print('ERR', e)
"; done
