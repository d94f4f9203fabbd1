#!/bin/bash
# round 2, GPU call Q: umma_fb2 with coalesced E1 mapping
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_bf16.py -m gpu -q -x -p no:cacheprovider --timeout 120 > $O/q_tests_bf16.log 2>&1; echo "bf16 tests rc=$?"; tail -3 $O/q_tests_bf16.log
STGCN_B200_LIB=$PWD/build/variants/tl.so timeout 120 python tools/fb2_cycles_probe.py > $O/q_fb2_cycles.txt 2>&1
cat $O/q_fb2_cycles.txt
b() { name=$1; shift; timeout 200 python bench.py --no-extras --no-cpu-baseline --steps 20 --warmup 5 "$@" > $O/q_bench_$name.json 2> $O/q_bench_$name.err; echo "bench $name rc=$?"; tail -2 $O/q_bench_$name.err; 
python -c "
import json
d=json.loads(open('$O/q_bench_$name.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'])
for k in d['top_kernels'][:30]: print('   ', round(k['ms_per_step']*1000,1), k['key'][:90])
"; }
b base
