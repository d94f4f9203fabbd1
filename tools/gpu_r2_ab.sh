#!/bin/bash
# round 2, GPU call AB: FB_FIRST_TC (P, Q of the first conv's backward from a tcgen05.mma instead of per-thread recompute), A/B
mkdir -p gpurun_out; O=gpurun_out
STGCN_FB0_TC=1 timeout 600 python -m pytest tests/test_gpu_bf16.py -m gpu -q -p no:cacheprovider --timeout 300 > $O/ab_tests_tc.log 2>&1; echo "bf16 tests (TC) rc=$?"; tail -3 $O/ab_tests_tc.log | cut -c1-300
timeout 600 python -m pytest tests/test_gpu_bf16.py -m gpu -q -p no:cacheprovider --timeout 300 -k "second_conv" > $O/ab_tests_fb2.log 2>&1; echo "fb2 tests rc=$?"; tail -2 $O/ab_tests_fb2.log | cut -c1-300
b() { name=$1; shift; timeout 200 python bench.py --no-extras --no-cpu-baseline --steps 20 --warmup 5 "$@" > $O/ab_bench_$name.json 2> $O/ab_bench_$name.err; echo "bench $name rc=$?"; 
python -c "
import json
d=json.loads(open('$O/ab_bench_$name.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'])
for k in d['top_kernels'][:6]: print('   ', round(k['ms_per_step']*1000,1), k['key'][:90])
"; }
b base
STGCN_FB0_TC=1 b tc
