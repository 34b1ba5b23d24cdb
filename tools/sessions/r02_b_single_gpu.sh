#!/bin/bash
# Round-2 single-GPU measurement session (run with gpurun from the repo root); outputs under gpurun_out/.
mkdir -p gpurun_out; export EZK_TRACE_CACHE=/tmp/ezk_cache
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/r9_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r9_pytest.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r9_bench.json 2> gpurun_out/r9_bench.err
for g in 28 14 7 4 2; do echo "EZK_LDE_GROUP=$g"; EZK_LDE_GROUP=$g python tools/profile_prove.py 20; done > gpurun_out/r9_lde_group.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'ntt_strided|ntt_final|constraint_kernel|hash_rows' --launch-skip 20 --launch-count 21 \
    -o gpurun_out/r9_full python tools/profile_prove.py 20 > gpurun_out/r9_ncu_full.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --launch-skip 70 --launch-count 75 --csv \
    --log-file gpurun_out/r9_traffic.csv python tools/profile_prove.py 20 > gpurun_out/r9_ncu_traffic.log 2>&1
python tools/prove_big.py 24 3 2 > gpurun_out/r9_big24.log 2>&1; cp profiles/r02_prove_2p24.json gpurun_out/ 2>/dev/null
tail -3 gpurun_out/r9_pytest.log; grep -h "EZK_LDE\|stages" gpurun_out/r9_lde_group.log; tail -2 gpurun_out/r9_big24.log; ls -la gpurun_out/r9_full.ncu-rep
