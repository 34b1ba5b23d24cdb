#!/bin/bash
# A/B of the CTA order of the NTT passes (tile-major vs column-major) on one B200: 2^20 (configs[2]) and 2^22 (configs[3])
# proofs, the precomputed-form inter-pass table under the new order, DRAM traffic per launch under the candidate order.
mkdir -p gpurun_out; export EZK_TRACE_CACHE=/tmp/ezk_cache
timeout 200 python tools/ntt_order_ab.py 20 2 3 > gpurun_out/rh_order_2p20.log 2>&1; tail -14 gpurun_out/rh_order_2p20.log
timeout 150 python tools/ntt_order_ab.py 22 3 2 -1,-1 0,-1 1,-1 3,-1 1,1 -1,-1 > gpurun_out/rh_order_2p22.log 2>&1; tail -8 gpurun_out/rh_order_2p22.log
cp encrypt_zkvm_b200/libezkvm.so /tmp/libezkvm_main.so; cp gpurun_scratch/libezkvm_prepass.so encrypt_zkvm_b200/libezkvm.so
timeout 100 python tools/ntt_order_ab.py 20 2 3 -1,-1 1,-1 2,-1 > gpurun_out/rh_order_2p20_prepass.log 2>&1; tail -5 gpurun_out/rh_order_2p20_prepass.log
cp /tmp/libezkvm_main.so encrypt_zkvm_b200/libezkvm.so
#EZK_NTT_ORDER=1 timeout 150 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --launch-skip 70 --launch-count 75 --csv \
#    --log-file gpurun_out/rh_traffic_order1.csv python tools/profile_prove.py 20 > gpurun_out/rh_ncu_traffic.log 2>&1
# (the traffic capture ran in the first attempt of this session)
