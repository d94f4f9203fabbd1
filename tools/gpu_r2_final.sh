#!/bin/bash
# round 2, final measurement pass on one B200: GPU suite, the driver's bench line (both arms), the other workloads, an ncu
# launch list, one `ncu --set full` capture of the dominant kernels, compute-sanitizer memcheck of smoke().
mkdir -p gpurun_out; O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 300 > $O/w_tests.log 2>&1; echo "tests rc=$?"; tail -5 $O/w_tests.log | cut -c1-300
(time timeout 600 python bench.py) > $O/w_bench.json 2> $O/w_bench.err; echo "bench default rc=$?"; tail -3 $O/w_bench.err | cut -c1-300
timeout 400 python bench.py --impl reference --steps 5 --warmup 3 > $O/w_ref.json 2> $O/w_ref.err; echo "ref rc=$?"; cut -c1-300 $O/w_ref.json | tail -1
b() { name=$1; shift; timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 20 --warmup 5 "$@" > $O/w_bench_$name.json 2> $O/w_bench_$name.err; echo "bench $name rc=$?"; tail -1 $O/w_bench_$name.err | cut -c1-200; }
b metrla --workload metrla
b pemsbay --workload pemsbay --batch 256
b syn2048 --workload syn2048 --batch 32 --steps 5 --warmup 3
b x3 --precision tf32x3
b drop05 --droprate 0.5
for f in $O/w_bench.json $O/w_bench_*.json; do echo $f; python -c "
import json,sys
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline_step']['frac'], d.get('gpu_launches'))
    for k in ('parity_mode','cfg4_pemsbay','cfg3_metrla','cuda_baseline','cpu_baseline'):
        if k in d and d[k]: print('   ', k, str(d[k])[:200])
    for k in d['top_kernels'][:8]: print('   ', round(k['ms_per_step']*1000,1), k['key'][:90])
except Exception as e: print('ERR', e)
"; done
timeout 250 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 200 --csv --log-file $O/w_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-profile --no-extras > $O/w_ncu1.log 2>&1; echo "ncu list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on \
  -k "regex:umma_fb2_kernel|umma_tap_kernel|umma_fb0_kernel|ln_bwd_sums_pg|umma_cheb_kernel|smallc1_conv_gate_fwd" -s 12 -c 26 -f -o $O/w_full \
  python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-profile --no-extras > $O/w_ncu2.log 2>&1; echo "ncu full rc=$?"
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $O/w_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -3 $O/w_memcheck.log
ls -la $O | grep " w_" | awk '{print $5, $9}'
