#!/bin/bash
# Check of the last host-side changes on one B200: the CTA-order test, the staged-upload tests (new copy pool), and a
# bench line without the CPU legs (e2e from pageable memory with the new pool).
mkdir -p gpurun_out
(time timeout 300 python -m pytest tests -m gpu -x -q -k "cta_order or staged or pageable or upload") > gpurun_out/rm_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/rm_pytest.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/rm_bench.json 2> gpurun_out/rm_bench.err; echo "bench rc=$?" >> gpurun_out/rm_bench.err
tail -4 gpurun_out/rm_pytest.log; tail -2 gpurun_out/rm_bench.err
python -c "
import json; d=json.load(open('gpurun_out/rm_bench.json')); print(d['value'], d['ms_per_step'], d['e2e'], d['checks'])"
