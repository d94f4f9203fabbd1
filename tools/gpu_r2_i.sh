#!/bin/bash
# round 2, GPU call I: division-free single-thread roles (RingPos) in tap / wgrad / gso kernels, store warp releases early
mkdir -p gpurun_out; O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -p no:cacheprovider --timeout 120 > $O/i_tests.log 2>&1; echo "tests rc=$?"; tail -4 $O/i_tests.log
KO_TAG=base timeout 120 python tools/ko_probe.py > $O/i_ko.txt 2>&1
for v in ko_se ko_sel ko_all; do
  KO_TAG=$v STGCN_B200_LIB=$PWD/build/variants/$v.so timeout 120 python tools/ko_probe.py >> $O/i_ko.txt 2>&1
done
cat $O/i_ko.txt
timeout 200 python bench.py --no-extras --no-cpu-baseline --steps 20 --warmup 5 > $O/i_bench_base.json 2> $O/i_bench_base.err; echo "bench rc=$?"; tail -2 $O/i_bench_base.err
python -c "
import json
d=json.loads(open('$O/i_bench_base.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'])
for k in d['top_kernels'][:40]: print('   ', round(k['ms_per_step']*1000,1), k['key'][:90])
"
