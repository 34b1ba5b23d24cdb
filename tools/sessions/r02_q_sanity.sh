#!/bin/bash
# Sanity of the final library of round 2 (after the launch-group rule moved into host/launch_groups.h): upload tests,
# smoke, one 2^20 proof from pageable memory compared with the digest every earlier run produced.
mkdir -p gpurun_out
(timeout 120 python -m pytest tests -m gpu -x -q -k "staged or pageable or upload or bookkeeping or cta_order or two_provers") > gpurun_out/rs_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/rs_pytest.log
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/rs_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/rs_smoke.log
timeout 60 python - > gpurun_out/rs_e2e.log 2>&1 <<'P'
import hashlib, json, os, sys, time
sys.path.insert(0, os.getcwd())
import encrypt_zkvm_b200 as ezk
prog, ex = ezk.synthetic_case(2, 20)
trace, ph, outs = ex.trace(), prog.hash(), ex.outputs()
with ezk.ExecutionProver(ezk.ProofOptions(), ph, outs, ezk.ServerKey()) as p:
    for _ in range(3): proof = p.prove(trace)
    w = []
    for _ in range(7):
        t0 = time.perf_counter(); proof = p.prove(trace); w.append((time.perf_counter() - t0) * 1e3)
    w.sort()
    p.verify(proof)
    d = hashlib.sha256(proof.to_bytes()).hexdigest()[:16]
    print(json.dumps({"median_ms": round(w[3], 3), "min_ms": round(w[0], 3), "sha256_16": d, "expected": "b436fe77d643852c", "ok": d == "b436fe77d643852c"}))
P
tail -3 gpurun_out/rs_pytest.log; tail -4 gpurun_out/rs_smoke.log; cat gpurun_out/rs_e2e.log
