#!/bin/bash
# round 2, GPU call N: tap kernel division-free item iterator (TapIter), uniform issuer prologue, per-shape issuer loops
mkdir -p gpurun_out; O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -p no:cacheprovider --timeout 120 > $O/n_tests.log 2>&1; echo "tests rc=$?"; tail -4 $O/n_tests.log
STGCN_TAP_TP2=1 timeout 600 python -m pytest tests/test_gpu_bf16.py -m gpu -q -x -p no:cacheprovider --timeout 120 > $O/n_tests_tp2.log 2>&1; echo "tests tp2 rc=$?"; tail -2 $O/n_tests_tp2.log
STGCN_B200_LIB=$PWD/build/variants/tl.so timeout 120 python tools/tap_cycles_probe.py > $O/n_cycles.txt 2>&1
echo "=========== TP2" >> $O/n_cycles.txt
STGCN_TAP_TP2=1 STGCN_B200_LIB=$PWD/build/variants/tl.so timeout 120 python tools/tap_cycles_probe.py >> $O/n_cycles.txt 2>&1
b() { name=$1; shift; timeout 200 python bench.py --no-extras --no-cpu-baseline --steps 20 --warmup 5 "$@" > $O/n_bench_$name.json 2> $O/n_bench_$name.err; echo "bench $name rc=$?"; tail -2 $O/n_bench_$name.err; 
python -c "
import json
d=json.loads(open('$O/n_bench_$name.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'])
for k in d['top_kernels'][:34]: print('   ', round(k['ms_per_step']*1000,1), k['key'][:90])
"; }
b base
STGCN_TAP_TP2=1 b tp2
