#!/bin/bash
# round 2, GPU call D (2 GPUs): full GPU suite incl. the 2-GPU NCCL test and the FB_GATE kernel, 1-GPU and 2-GPU bench lines.
mkdir -p gpurun_out; O=gpurun_out
nvidia-smi -L > $O/d_smi.txt
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 300 > $O/d_tests.log 2>&1; echo "tests rc=$?" | tee -a $O/d_tests.log
tail -12 $O/d_tests.log
timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 20 --warmup 5 > $O/d_bench_1gpu.json 2> $O/d_bench_1gpu.err; echo "bench 1gpu rc=$?"
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > $O/d_bench_2gpu.json 2> $O/d_bench_2gpu.err; echo "bench 2gpu rc=$?"; tail -5 $O/d_bench_2gpu.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-extras --reduce-after > $O/d_bench_2gpu_after.json 2> $O/d_bench_2gpu_after.err; echo "bench 2gpu reduce-after rc=$?"; tail -3 $O/d_bench_2gpu_after.err
STGCN_NO_SIDE_STREAMS=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 --no-extras > $O/d_bench_2gpu_noside.json 2> $O/d_bench_2gpu_noside.err; echo "bench 2gpu no-side rc=$?"
for f in $O/d_bench*.json; do echo $f; python -c "
import json,sys
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['config'].get('grad_allreduce'), d['clocks'])
    for k in ('parity_mode','cfg4_pemsbay','cfg3_metrla'):
        if k in d: print('   ', k, d[k])
    for k in d['top_kernels'][:6]: print('   ', round(k['ms_per_step']*1000,1), k['key'][:90])
except Exception as e: print('ERR', e)
"; done
