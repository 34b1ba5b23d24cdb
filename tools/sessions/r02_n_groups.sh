#!/bin/bash
# Same-box A/B of the host-trace pipeline: previous commit (fixed 2-column groups) vs launch groups that grow while the
# upload runs ahead; the upload tests with the new library first.
mkdir -p gpurun_out
(time timeout 300 python -m pytest tests -m gpu -x -q -k "staged or pageable or upload or bookkeeping or errors or shape") > gpurun_out/rn_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/rn_pytest.log
tail -4 gpurun_out/rn_pytest.log
cp encrypt_zkvm_b200/libezkvm.so /tmp/libezkvm_main.so
for lib in head main head main; do
  if [ $lib = head ]; then cp gpurun_scratch/libezkvm_head.so encrypt_zkvm_b200/libezkvm.so; else cp /tmp/libezkvm_main.so encrypt_zkvm_b200/libezkvm.so; fi
  timeout 200 python tools/pageable_e2e.py 20 2 7 > gpurun_out/rn_pageable_$lib.json 2> gpurun_out/rn_pageable_$lib.err
  python - <<P
import json
d=json.load(open("gpurun_out/rn_pageable_$lib.json"))
print("$lib", {k:(round(v["ms_per_proof"],2), round(v["min_ms"],2)) for k,v in d.items() if isinstance(v,dict) and "ms_per_proof" in v}, d["identical_bytes"])
P
done
cp /tmp/libezkvm_main.so encrypt_zkvm_b200/libezkvm.so
