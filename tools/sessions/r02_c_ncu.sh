#!/bin/bash
# Round-2 ncu session: full capture of the hot kernels of one 2^20 proof, summarised ON the box (the .ncu-rep itself is
# larger than what gpurun copies back), DRAM traffic of every launch of a proof, and a bench line.
mkdir -p gpurun_out; export EZK_TRACE_CACHE=/tmp/ezk_cache
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r9_bench.json 2> gpurun_out/r9_bench.err
ncu --set full --clock-control none --import-source on -k regex:'ntt_strided|ntt_final|constraint_kernel|hash_rows' --launch-skip 20 --launch-count 21 \
    -o /tmp/r9_full python tools/profile_prove.py 20 > gpurun_out/r9_ncu_full.log 2>&1
python tools/ncu_summary.py full /tmp/r9_full.ncu-rep gpurun_out/r02_ncu_full > /dev/null 2> gpurun_out/r9_ncu_summary.err
cp profiles/roofline_traffic.json gpurun_out/roofline_traffic_from_full.json
ncu -i /tmp/r9_full.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_raw.csv 2>/dev/null
ncu -i /tmp/r9_full.ncu-rep --page source --csv -k regex:ntt_strided -c 1 > gpurun_out/r02_ncu_source_strided.csv 2>/dev/null
ncu -i /tmp/r9_full.ncu-rep --page source --csv -k regex:ntt_final -c 1 > gpurun_out/r02_ncu_source_final.csv 2>/dev/null
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --launch-skip 70 --launch-count 75 --csv \
    --log-file gpurun_out/r9_traffic.csv python tools/profile_prove.py 20 > gpurun_out/r9_ncu_traffic.log 2>&1
du -sh gpurun_out; ls -la gpurun_out | head -30
