#!/bin/bash
# Same-box A/B of the staged upload from pageable memory: library of the previous commit (fork-join copy in equal parts,
# 4 threads) vs the dynamic-piece copy pool (8 threads by default).
mkdir -p gpurun_out
cp encrypt_zkvm_b200/libezkvm.so /tmp/libezkvm_main.so; cp gpurun_scratch/libezkvm_head.so encrypt_zkvm_b200/libezkvm.so
timeout 200 python tools/pageable_e2e.py 20 2 7 > gpurun_out/rl_pageable_head.json 2> gpurun_out/rl_pageable_head.err
cp /tmp/libezkvm_main.so encrypt_zkvm_b200/libezkvm.so
timeout 200 python tools/pageable_e2e.py 20 2 7 > gpurun_out/rl_pageable_new.json 2> gpurun_out/rl_pageable_new.err
cp encrypt_zkvm_b200/libezkvm.so /tmp/libezkvm_main.so; cp gpurun_scratch/libezkvm_head.so encrypt_zkvm_b200/libezkvm.so
timeout 200 python tools/pageable_e2e.py 20 2 7 > gpurun_out/rl_pageable_head2.json 2> gpurun_out/rl_pageable_head2.err
cp /tmp/libezkvm_main.so encrypt_zkvm_b200/libezkvm.so
python - <<'P'
import json
for f in ("head","new","head2"):
    try:
        d=json.load(open(f"gpurun_out/rl_pageable_{f}.json"))
        print(f, {k:(round(v["ms_per_proof"],2), round(v["min_ms"],2)) for k,v in d.items() if isinstance(v,dict) and "ms_per_proof" in v}, d["identical_bytes"])
    except Exception as e: print(f, "failed", e)
P
nproc
