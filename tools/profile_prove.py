"""One warm-up proof and one measured proof at 2^log_n rows (for ncu: skip the warm-up launches with --launch-skip).

    python tools/profile_prove.py [log_n] [kind]      # prints device ms, stage ms and per-kernel ms of the measured proof
"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
import encrypt_zkvm_b200 as ezk

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
kind = int(sys.argv[2]) if len(sys.argv) > 2 else 2
prog, ex = ezk.synthetic_case(kind, log_n)
trace = ex.trace()
dev = torch.from_numpy(trace.view(np.int64)).to("cuda:0")
torch.cuda.synchronize()
p = ezk.ExecutionProver(ezk.ProofOptions(), prog.hash(), ex.outputs(), ezk.ServerKey())
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
