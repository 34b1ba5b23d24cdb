#!/bin/bash
# Second final single-GPU session of round 2 (after the tile-major CTA order and the one-load coset factor): GPU tests,
# smoke, the bench line with the driver's flags, ncu launch list, ncu full capture of the shipped kernels (summarised on
# the box), per-instruction counters of the LDE's strided pass.
mkdir -p gpurun_out; export EZK_TRACE_CACHE=/tmp/ezk_cache
(time timeout 600 python -m pytest tests -m gpu -x -q) > gpurun_out/rk_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/rk_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/rk_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/rk_smoke.log
(time python bench.py --steps 20 --warmup 5 > gpurun_out/rk_bench.json) 2> gpurun_out/rk_bench.err; echo "bench rc=$?" >> gpurun_out/rk_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/rk_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/rk_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'ntt_strided|ntt_final|constraint_kernel|hash_rows|merkle_subtree' --launch-skip 22 --launch-count 23 \
    -o /tmp/rk_full python tools/profile_prove.py 20 > gpurun_out/rk_ncu_full.log 2>&1
python tools/ncu_summary.py full /tmp/rk_full.ncu-rep gpurun_out/r02_ncu_full_final2 > /dev/null 2> gpurun_out/rk_ncu_summary.err
ncu -i /tmp/rk_full.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_final2_raw.csv 2>/dev/null
ncu -i /tmp/rk_full.ncu-rep --page source --csv -k regex:ntt_strided -s 1 -c 1 > gpurun_out/rk_source_strided_lde.csv 2>/dev/null
du -sh gpurun_out; tail -3 gpurun_out/rk_pytest.log; tail -4 gpurun_out/rk_smoke.log; tail -4 gpurun_out/rk_bench.err; cut -c1-700 gpurun_out/rk_bench.json
