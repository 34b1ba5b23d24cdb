#!/bin/bash
# A/B of the NTT build options on one B200: in-tile twiddles in precomputed form (0 off / 1 both passes / 2 final pass)
# x launch shape of the final pass (0 = 256 threads x 3 CTAs/SM, 2 = 256 x 2).  Kernel times of one 2^20 proof.
mkdir -p gpurun_out; export EZK_TRACE_CACHE=/tmp/ezk_cache
cp encrypt_zkvm_b200/libezkvm.so /tmp/libezkvm_pre2.so; cp gpurun_scratch/libezkvm_pre0.so gpurun_scratch/libezkvm_pre1.so /tmp/
for lib in pre2 pre0 pre1; do
  cp /tmp/libezkvm_$lib.so encrypt_zkvm_b200/libezkvm.so
  for fv in 0 2; do
    echo "== lib=$lib EZK_NTT_FINAL_VARIANT=$fv"
    EZK_NTT_FINAL_VARIANT=$fv python tools/profile_prove.py 20 | grep -v "^log_n" | sed 's/, .constraints.*//'
    EZK_NTT_FINAL_VARIANT=$fv python tools/profile_prove.py 20 | grep "^kernels" | sed 's/, .constraints.*//'
  done
done > gpurun_out/r11_ntt_ab.log 2>&1
cp /tmp/libezkvm_pre2.so encrypt_zkvm_b200/libezkvm.so
cat gpurun_out/r11_ntt_ab.log
