"""Proof-level pipelining: P provers on ONE GPU, one host thread each, proving independent traces.

    python tools/concurrent_provers.py [log_n] [kind] [provers ...]      # default 20 2 1 2 3

A prover owns its streams, workspace and tables, and the library holds no lock across provers (ctypes releases the GIL
during the call), so the latency-bound stretches of one proof — Merkle tops, FRI tail, the host's transcript round
trips, query gathering — are filled with the kernels of another.  Prints aggregate proofs/s per prover count; every
proof is checked to be byte-identical to the one a single prover returns.
"""
import os
import pickle
import sys
import threading
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
import encrypt_zkvm_b200 as ezk

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
kind = int(sys.argv[2]) if len(sys.argv) > 2 else 2
counts = [int(x) for x in sys.argv[3:]] or [1, 2, 3]
steps = 8 if log_n >= 20 else 40

cache = os.environ.get("EZK_TRACE_CACHE")
cache_file = Path(cache) / f"case_{kind}_{log_n}.pkl" if cache else None
if cache_file and cache_file.exists():
    trace, program_hash, outputs = pickle.loads(cache_file.read_bytes())
else:
    prog, ex = ezk.synthetic_case(kind, log_n)
    trace, program_hash, outputs = ex.trace(), prog.hash(), ex.outputs()
    if cache_file:
        cache_file.parent.mkdir(parents=True, exist_ok=True)
        cache_file.write_bytes(pickle.dumps((trace, program_hash, outputs), protocol=4))
n = 1 << log_n
dev = torch.from_numpy(trace.view(np.int64)).to("cuda:0")
torch.cuda.synchronize()

reference = None
for P in counts:
    provers = [ezk.ExecutionProver(ezk.ProofOptions(), program_hash, outputs, ezk.ServerKey()) for _ in range(P)]
    results = [None] * P

    def work(k, reps):
        out = None
        for _ in range(reps):
            out = provers[k].prove_device(dev.data_ptr(), n).to_bytes()
        results[k] = out

    for reps in (2, steps):  # warm-up, then the timed run
        threads = [threading.Thread(target=work, args=(k, reps)) for k in range(P)]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    if reference is None:
        reference = results[0]
    same = all(r == reference for r in results)
    print(f"log_n={log_n} provers={P} proofs={P * steps} wall_s={dt:.4f} proofs_per_s={P * steps / dt:.2f} "
          f"ms_per_proof={dt * 1e3 / (P * steps):.3f} identical={same}", flush=True)
    for p in provers:
        p.close()
