#!/bin/bash
# round 2, GPU call O: umma_fb2 (LayerNorm backward + gate backward + tc2 data/weight gradients in one kernel)
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_bf16.py -m gpu -q -x -p no:cacheprovider --timeout 120 > $O/o_tests_bf16.log 2>&1; echo "bf16 tests rc=$?"; tail -15 $O/o_tests_bf16.log
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 120 > $O/o_tests.log 2>&1; echo "tests rc=$?"; tail -6 $O/o_tests.log
b() { name=$1; shift; timeout 200 python bench.py --no-extras --no-cpu-baseline --steps 20 --warmup 5 "$@" > $O/o_bench_$name.json 2> $O/o_bench_$name.err; echo "bench $name rc=$?"; tail -2 $O/o_bench_$name.err; 
python -c "
import json
d=json.loads(open('$O/o_bench_$name.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'])
for k in d['top_kernels'][:30]: print('   ', round(k['ms_per_step']*1000,1), k['key'][:90])
"; }
b base
STGCN_NO_FB2=1 b nofb2
