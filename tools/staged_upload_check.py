"""Torch-free check of the staged upload (default for pageable memory; EZK_STAGED_UPLOAD=0 = plain copy): the proof of a pageable host trace must have
the same bytes with and without it, at sizes with one chunk and with many chunks per column.

    python tools/staged_upload_check.py [log_n ...]
"""
import os
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import encrypt_zkvm_b200 as ezk

os.environ.setdefault("EZK_STAGE_SLOT_KB", "64")
sizes = [int(a) for a in sys.argv[1:]] or [7, 12, 16]
for log_n in sizes:
    prog, ex = ezk.synthetic_case(2, log_n)
    trace, program_hash, outputs = ex.trace(), prog.hash(), ex.outputs()
    with ezk.ExecutionProver(ezk.ProofOptions(), program_hash, outputs, ezk.ServerKey()) as p:
        res = {}
        for mode in ("0", "1"):
            os.environ["EZK_STAGED_UPLOAD"] = mode
            p.prove(trace)
            t0 = time.perf_counter()
            proof = p.prove(trace).to_bytes()
            res[mode] = (proof, (time.perf_counter() - t0) * 1e3, p.stage_times_ms()["upload"])
        same = res["0"][0] == res["1"][0]
        print(f"log_n={log_n} same_bytes={same} plain: e2e {res['0'][1]:.2f} ms upload {res['0'][2]:.2f} ms | "
              f"staged: e2e {res['1'][1]:.2f} ms upload {res['1'][2]:.2f} ms", flush=True)
        if not same:
            sys.exit(1)
print("ok", flush=True)
