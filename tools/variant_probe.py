"""Proof bytes and device stage times of ONE build of the library (the default, or the build variant named by the
EZKVM_LIB environment variable - see VARIANTS in encrypt_zkvm_b200/build.py).  Torch-free; prints one JSON line.
bench.py runs it once per library in subprocesses and reports, as the informational `variants` key, whether a
variant returns the same proof bytes as the default and what each stage costs.

    [EZKVM_LIB=encrypt_zkvm_b200/libezkvm_pretw.so] python tools/variant_probe.py [log_n] [kind] [steps] [device]
"""
import hashlib
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import encrypt_zkvm_b200 as ezk
from encrypt_zkvm_b200 import _lib

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
kind = int(sys.argv[2]) if len(sys.argv) > 2 else 2
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 7
device = int(sys.argv[4]) if len(sys.argv) > 4 else 0
prog, ex = ezk.synthetic_case(kind, log_n)
trace, program_hash, outputs = ex.trace(), prog.hash(), ex.outputs()
samples = []
with ezk.ExecutionProver(ezk.ProofOptions(), program_hash, outputs, ezk.ServerKey(), device=device) as p:
    for _ in range(3):
        proof = p.prove(trace)
    for _ in range(steps):
        proof = p.prove(trace)
        samples.append(p.stage_times_ms())
    p.verify(proof)
med = {k: sorted(s[k] for s in samples)[len(samples) // 2] for k in samples[0]}
print(json.dumps({"lib": Path(_lib.LIB_PATH).name, "log_n": log_n, "steps": steps,
                  "proof_sha256": hashlib.sha256(proof.to_bytes()).hexdigest(), "verified": True, "stage_ms": med,
                  "device_ms_without_upload": sum(v for k, v in med.items() if k != "upload")}), flush=True)
