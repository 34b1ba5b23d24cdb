"""One warm-up proof and one measured proof at 2^log_n rows (for ncu: skip the warm-up launches with --launch-skip).

    python tools/profile_prove.py [log_n] [kind]      # prints device ms, stage ms and per-kernel ms of the measured proof

EZK_TRACE_CACHE=dir keeps the generated trace (the host VM needs ~30 s for 2^20 rows) for the next invocation.
"""
import os
import pickle
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
import encrypt_zkvm_b200 as ezk

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
kind = int(sys.argv[2]) if len(sys.argv) > 2 else 2
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
dev = torch.from_numpy(trace.view(np.int64)).to("cuda:0")
torch.cuda.synchronize()
p = ezk.ExecutionProver(ezk.ProofOptions(), program_hash, outputs, ezk.ServerKey())
before = ezk.kernel_launch_count()
p.prove_device(dev.data_ptr(), 1 << log_n)
per_proof = ezk.kernel_launch_count() - before
ezk.profile_enable(True)
ezk.profile_reset()
p.timer_start()
proof = p.prove_device(dev.data_ptr(), 1 << log_n)
ms = p.timer_stop()
prof = ezk.profile_read()
ezk.profile_enable(False)
print(f"log_n={log_n} kernels_per_proof={per_proof} proof_bytes={len(proof)} device_ms={ms:.3f}")
print("stages:", {k: round(v, 3) for k, v in p.stage_times_ms().items()})
print("kernels:", {k: round(v["ms"], 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])})
p.close()
