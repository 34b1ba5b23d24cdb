#!/bin/bash
# Final single-GPU session of round 2: GPU tests, the bench line with the driver's flags, reference arm, ncu launch list,
# ncu full capture of the shipped kernels (summarised on the box), per-launch DRAM traffic.
mkdir -p gpurun_out; export EZK_TRACE_CACHE=/tmp/ezk_cache
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/rf_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/rf_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/rf_bench.json 2> gpurun_out/rf_bench.err; echo "bench rc=$?" >> gpurun_out/rf_bench.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/rf_ref.json 2> gpurun_out/rf_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/rf_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/rf_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'ntt_strided|ntt_final|constraint_kernel|hash_rows|merkle_subtree' --launch-skip 22 --launch-count 23 \
    -o /tmp/rf_full python tools/profile_prove.py 20 > gpurun_out/rf_ncu_full.log 2>&1
python tools/ncu_summary.py full /tmp/rf_full.ncu-rep gpurun_out/r02_ncu_full_final > /dev/null 2> gpurun_out/rf_ncu_summary.err
ncu -i /tmp/rf_full.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_final_raw.csv 2>/dev/null
ncu -i /tmp/rf_full.ncu-rep --page source --csv -k regex:ntt_strided -s 1 -c 1 > gpurun_out/rf_source_strided_lde.csv 2>/dev/null
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --launch-skip 70 --launch-count 75 --csv \
    --log-file gpurun_out/rf_traffic.csv python tools/profile_prove.py 20 > gpurun_out/rf_ncu_traffic.log 2>&1
du -sh gpurun_out; tail -3 gpurun_out/rf_pytest.log; cut -c1-600 gpurun_out/rf_bench.json
