#!/bin/bash
# e2e (pageable / pinned) and device-resident bench lines for the in-tile twiddle forms 0 and 2, same box.
mkdir -p gpurun_out
cp encrypt_zkvm_b200/libezkvm.so /tmp/libezkvm_pre2.so; cp gpurun_scratch/libezkvm_pre0.so /tmp/
for lib in pre2 pre0 pre2 pre0; do
  cp /tmp/libezkvm_$lib.so encrypt_zkvm_b200/libezkvm.so
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('$lib', 'device', round(d['ms_per_step'], 3), 'e2e_pageable', round(d['e2e']['ms_per_step'], 3), 'e2e_pinned', round(d['e2e']['pinned']['ms_per_step'], 3), 'pipelined', round(d['pipelined']['proofs_per_s'], 2), 'lde', d['stages']['trace_lde']['ms'], 'strided', d['kernels']['ntt_strided_pass']['ms'], 'final', d['kernels']['ntt_final_pass']['ms'])"
done > gpurun_out/r12_pre_e2e.log 2>&1
cp /tmp/libezkvm_pre2.so encrypt_zkvm_b200/libezkvm.so
cat gpurun_out/r12_pre_e2e.log
