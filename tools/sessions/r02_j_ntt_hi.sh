#!/bin/bash
# Same-box A/B: library of the previous commit vs the coset load factor as one load from the upper table level.
mkdir -p gpurun_out; export EZK_TRACE_CACHE=/tmp/ezk_cache
cp encrypt_zkvm_b200/libezkvm.so /tmp/libezkvm_main.so; cp gpurun_scratch/libezkvm_head.so encrypt_zkvm_b200/libezkvm.so
timeout 120 python tools/ntt_order_ab.py 20 2 3 - - > gpurun_out/rj_head_2p20.log 2>&1; tail -3 gpurun_out/rj_head_2p20.log
cp /tmp/libezkvm_main.so encrypt_zkvm_b200/libezkvm.so
timeout 120 python tools/ntt_order_ab.py 20 2 3 - - - > gpurun_out/rj_new_2p20.log 2>&1; tail -4 gpurun_out/rj_new_2p20.log
timeout 120 python tools/ntt_order_ab.py 22 3 2 - - > gpurun_out/rj_new_2p22.log 2>&1; tail -3 gpurun_out/rj_new_2p22.log
timeout 100 python tools/ntt_order_ab.py 16 1 5 - - > gpurun_out/rj_new_2p16.log 2>&1; tail -3 gpurun_out/rj_new_2p16.log
