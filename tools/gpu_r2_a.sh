#!/bin/bash
# round 2, GPU call A: full GPU test suite (new tf32x3 mode, large-N bf16, graph/dropout/micro-stream tests), default bench
# line with extras, candidate A/Bs (LN group kernel, first-layer prefetch, micro-streams), never-run variants, sanitizer.
mkdir -p gpurun_out; O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/a_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 180 > $O/a_tests.log 2>&1; echo "tests rc=$?" | tee -a $O/a_tests.log
tail -40 $O/a_tests.log
timeout 400 python bench.py > $O/a_bench.json 2> $O/a_bench.err; echo "bench rc=$?"
b() { name=$1; shift; timeout 240 env "$@" python bench.py --no-extras --no-cpu-baseline --steps 20 --warmup 5 ${BARGS} > $O/a_bench_$name.json 2> $O/a_bench_$name.err; echo "bench $name rc=$?"; }
BARGS="" b base X=1
BARGS="" b lngroup STGCN_LN_GROUP=1
BARGS="" b prefetch STGCN_SMALLC1_PREFETCH=1
BARGS="--micro-streams 2" b micro2 X=1
BARGS="--precision tf32x3" b x3 X=1
BARGS="--precision fp32" b fp32 X=1
BARGS="--no-graph" b nograph X=1
BARGS="--droprate 0.5" b drop05 X=1
STGCN_LN_GROUP=1 timeout 300 python -m pytest tests/test_gpu_bf16.py -q -p no:cacheprovider --timeout 120 -k "model or graphed" > $O/a_tests_lngroup.log 2>&1; echo "lngroup tests rc=$?"; tail -3 $O/a_tests_lngroup.log
STGCN_SMALLC1_PREFETCH=1 timeout 300 python -m pytest tests/test_gpu_bf16.py tests/test_gpu_parity.py -q -p no:cacheprovider --timeout 120 -k "model or temporal" > $O/a_tests_prefetch.log 2>&1; echo "prefetch tests rc=$?"; tail -3 $O/a_tests_prefetch.log
STGCN_GSO_KTILED=1 timeout 300 python -m pytest tests/test_gpu_bf16.py -q -p no:cacheprovider --timeout 120 -k "graph_conv" > $O/a_tests_ktiled.log 2>&1; echo "ktiled tests rc=$?"; tail -3 $O/a_tests_ktiled.log
timeout 500 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > $O/a_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -5 $O/a_memcheck.log
head -c 1500 $O/a_bench.json; echo; for f in $O/a_bench_*.json; do echo $f; python -c "
import json,sys
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'])
except Exception as e: print('ERR', e)
"; done
