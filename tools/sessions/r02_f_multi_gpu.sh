#!/bin/bash
# Multi-GPU session (gpurun --gpus N): byte identity + timing of the sharded proof at 2^20 and 2^22, then the bench line.
N=${1:-8}; mkdir -p gpurun_out
timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 tools/sharded_check.py 20 22 > gpurun_out/r13_sharded$N.log 2>&1
grep log_n gpurun_out/r13_sharded$N.log
timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r13_bench$N.json 2> gpurun_out/r13_bench$N.err
tail -2 gpurun_out/r13_bench$N.err
python -c "
import json; d=json.load(open('gpurun_out/r13_bench$N.json')); print(d['value'], d['ms_per_step'], d.get('speedup_vs_one_gpu'), d['single_gpu_same_config'], d['e2e'], d['kernels'], d['stages'], d['checks'])"
