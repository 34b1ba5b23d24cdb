"""End-to-end proofs from a PAGEABLE host trace (what a Rust `Vec<BaseElement>` column is), with the plain upload
and with the staged upload (the default for unregistered memory; EZK_STAGED_UPLOAD=0 disables it;
csrc/host/copy_pool.h), for 1 / 2 / 4 / 8 / the default number of copy threads.  Torch-free; prints one JSON line and writes it to
profiles/r02_pageable_e2e_2p<log_n>.json.

    python tools/pageable_e2e.py [log_n] [kind] [steps] [device]
"""
import json
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import encrypt_zkvm_b200 as ezk

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
kind = int(sys.argv[2]) if len(sys.argv) > 2 else 2
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
device = int(sys.argv[4]) if len(sys.argv) > 4 else 0
t0 = time.perf_counter()
prog, ex = ezk.synthetic_case(kind, log_n)
trace, program_hash, outputs = ex.trace(), prog.hash(), ex.outputs()  # numpy array: pageable memory
vm_s = time.perf_counter() - t0
res = {}


def measure(staged: bool):
    # a prover of its own per configuration: the copy pool is created on the first staged upload
    os.environ["EZK_STAGED_UPLOAD"] = "1" if staged else "0"
    with ezk.ExecutionProver(ezk.ProofOptions(), program_hash, outputs, ezk.ServerKey(), device=device) as p:
        for _ in range(2):
            p.prove(trace)
        wall = []
        for _ in range(steps):
            t0 = time.perf_counter()
            proof = p.prove(trace).to_bytes()
            wall.append((time.perf_counter() - t0) * 1e3)
        return {"ms_per_proof": sorted(wall)[len(wall) // 2], "min_ms": min(wall),
                "ms_before_first_launch": p.stage_times_ms()["upload"]}, proof


res["plain"], want = measure(False)
same = True
for threads in (0, 1, 2, 4, 8):  # 0 = the library's default (min(8, cores / 2))
    if threads:
        os.environ["EZK_STAGE_THREADS"] = str(threads)
    res[f"staged_{threads}_threads" if threads else "staged"], got = measure(True)
    same = same and got == want

# what INTEGRATION.md recommends to callers that prove many traces from the same buffers: page-lock them once
# (cudaHostRegister) and use the plain upload.  Cost of the registration and the per-proof time after it.
try:
    import ctypes
    rt = ctypes.CDLL("libcudart.so.12")
    rt.cudaHostRegister.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint]
    rt.cudaHostUnregister.argtypes = [ctypes.c_void_p]
    rt.cudaSetDevice(device)
    t0 = time.perf_counter()
    rc = rt.cudaHostRegister(trace.ctypes.data, trace.nbytes, 0)
    reg_ms = (time.perf_counter() - t0) * 1e3
    if rc == 0:
        res["registered"], got = measure(False)
        res["registered"]["cudaHostRegister_ms"] = reg_ms
        same = same and got == want
        rt.cudaHostUnregister(trace.ctypes.data)
    else:
        res["registered"] = {"error": f"cudaHostRegister returned {rc}"}
except Exception as e:  # optional datum
    res["registered"] = {"error": f"{type(e).__name__}: {e}"[:200]}
line = {"log_n": log_n, "steps": steps, "host_vm_s": vm_s, "identical_bytes": same, **res}
(Path(__file__).resolve().parent.parent / "profiles" / f"r02_pageable_e2e_2p{log_n}.json").write_text(json.dumps(line, indent=1) + "\n")
print(json.dumps(line), flush=True)
