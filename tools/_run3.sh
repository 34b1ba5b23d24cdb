cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
export EZK_TRACE_CACHE=/tmp/ezk_traces
O=gpurun_out/r01o_variants.log
: > $O
run() { echo "== $*" >> $O; env "$@" python tools/profile_prove.py 20 2>&1 | tail -3 >> $O; }
run A=default
run EZK_CONSTRAINT_VARIANT=0
run EZK_CONSTRAINT_VARIANT=1
run EZK_CONSTRAINT_VARIANT=3
run EZK_NTT_VARIANT=1
run EZK_NTT_VARIANT=2
run EZK_NTT_BIG_TABLE_MB=0
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "survives" >> $O 2>&1
# ncu: OOD evaluation kernels of the measured proof
ncu --set full --clock-control none --import-source on -k regex:"eval_|power_table" -o gpurun_out/r01o_eval python tools/profile_prove.py 20 > gpurun_out/r01o_ncu.log 2>&1
cat $O
