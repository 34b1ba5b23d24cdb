#!/bin/bash
# A/B of cp.async staging in the NTT passes (EZK_NTT_STAGE bits: 1 strided / 2 final first-step inputs of a thread's later
# groups; 4 strided / 8 final output factors of the last step) on one B200; the library of the previous commit first.
mkdir -p gpurun_out; export EZK_TRACE_CACHE=/tmp/ezk_cache
cp encrypt_zkvm_b200/libezkvm.so /tmp/libezkvm_main.so; cp gpurun_scratch/libezkvm_head.so encrypt_zkvm_b200/libezkvm.so
timeout 120 python tools/ntt_order_ab.py 20 2 3 - - > gpurun_out/ri_stage_2p20_head.log 2>&1; tail -3 gpurun_out/ri_stage_2p20_head.log
cp /tmp/libezkvm_main.so encrypt_zkvm_b200/libezkvm.so
timeout 200 python tools/ntt_order_ab.py 20 2 3 EZK_NTT_STAGE=0 EZK_NTT_STAGE=1 EZK_NTT_STAGE=2 EZK_NTT_STAGE=3 EZK_NTT_STAGE=4 EZK_NTT_STAGE=8 EZK_NTT_STAGE=5 EZK_NTT_STAGE=7 EZK_NTT_STAGE=15 EZK_NTT_STAGE=0 > gpurun_out/ri_stage_2p20.log 2>&1; tail -12 gpurun_out/ri_stage_2p20.log
timeout 150 python tools/ntt_order_ab.py 22 3 2 EZK_NTT_STAGE=0 EZK_NTT_STAGE=3 EZK_NTT_STAGE=7 EZK_NTT_STAGE=15 EZK_NTT_STAGE=0 > gpurun_out/ri_stage_2p22.log 2>&1; tail -6 gpurun_out/ri_stage_2p22.log
