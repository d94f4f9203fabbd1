#!/bin/bash
# round 2, last pass on one B200 (final code state): GPU suite, the driver's bench line (both arms), the other workloads,
# ncu launch list + --set full capture of the fused backward kernels, memcheck of smoke().
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 300 > $O/ac_tests.log 2>&1; echo "tests rc=$?"; tail -3 $O/ac_tests.log | cut -c1-300
(time timeout 600 python bench.py) > $O/ac_bench.json 2> $O/ac_bench.err; echo "bench default rc=$?"; tail -3 $O/ac_bench.err | cut -c1-200
timeout 400 python bench.py --impl reference --steps 5 --warmup 3 > $O/ac_ref.json 2> $O/ac_ref.err; echo "ref rc=$?"; cut -c1-200 $O/ac_ref.json | tail -1
b() { name=$1; shift; timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 20 --warmup 5 "$@" > $O/ac_bench_$name.json 2> $O/ac_bench_$name.err; echo "bench $name rc=$?"; }
b metrla --workload metrla
b pemsbay --workload pemsbay --batch 256
b syn2048 --workload syn2048 --batch 32 --steps 5 --warmup 3
b drop05 --droprate 0.5
for f in $O/ac_bench.json $O/ac_bench_*.json; do echo $f; python -c "
import json
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline_step']['frac'], d.get('gpu_launches'), d['roofline']['kernel'], round(d['roofline']['frac'],3), d['roofline'].get('traffic'))
    for k in d['top_kernels'][:6]: print('   ', round(k['ms_per_step']*1000,1), k['key'][:90])
except Exception as e: print('ERR', e)
"; done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 200 --csv --log-file $O/ac_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-profile --no-extras > $O/ac_ncu1.log 2>&1; echo "ncu list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:umma_fb2_kernel|umma_fb0_kernel|ln_bwd_sums_pg" -s 4 -c 8 -f -o $O/ac_full \
  python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-profile --no-extras > $O/ac_ncu2.log 2>&1; echo "ncu full rc=$?"
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $O/ac_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -2 $O/ac_memcheck.log
ls -la $O | grep " ac_" | awk '{print $5, $9}'
