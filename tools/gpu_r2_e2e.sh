#!/bin/bash
# round 2: end-to-end leg with the double-buffered input pipeline (HostBatchPrefetcher) vs the copy-in-front path
mkdir -p gpurun_out; O=gpurun_out
timeout 120 python bench.py --no-cpu-baseline --no-extras --no-profile --steps 20 --warmup 5 > $O/af_bench_pf.json 2> $O/af_bench_pf.err; echo "rc=$?"; tail -2 $O/af_bench_pf.err
STGCN_BENCH_NO_PREFETCH=1 timeout 120 python bench.py --no-cpu-baseline --no-extras --no-profile --steps 20 --warmup 5 > $O/af_bench_nopf.json 2> $O/af_bench_nopf.err; echo "rc=$?"
for f in $O/af_bench_pf.json $O/af_bench_nopf.json; do python -c "
import json
d=json.loads(open('$f').read().strip().splitlines()[-1]); print(round(d['value']), d['ms_per_step'], d['e2e'])
"; done
