#!/bin/bash
# round 2, final 2-GPU pass: NCCL gradient-parity test on the product layers, 1 -> 2 GPU bench lines
mkdir -p gpurun_out; O=gpurun_out
nvidia-smi -L > $O/x_smi.txt
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q -p no:cacheprovider --timeout 300 > $O/x_tests.log 2>&1; echo "dist tests rc=$?"; tail -3 $O/x_tests.log | cut -c1-300
timeout 300 python bench.py --no-extras --no-cpu-baseline --steps 20 --warmup 5 > $O/x_bench_1gpu.json 2> $O/x_bench_1gpu.err; echo "bench 1gpu rc=$?"
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > $O/x_bench_2gpu.json 2> $O/x_bench_2gpu.err; echo "bench 2gpu rc=$?"; tail -3 $O/x_bench_2gpu.err | cut -c1-300
for f in $O/x_bench*.json; do echo $f; python -c "
import json
try:
    d=json.loads(open('$f').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['config'].get('grad_allreduce'), d['clocks'])
    for k in ('cfg4_pemsbay',):
        if k in d and d[k]: print('   ', k, str(d[k])[:300])
except Exception as e: print('ERR', e)
"; done
