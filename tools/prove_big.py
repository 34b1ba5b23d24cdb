"""A full proof at a large trace length on ONE GPU (BASELINE.json configs[4], upper end: 2^24 rows x 28 columns):
device-resident trace, product verifier and oracle verifier on the result, device time and peak device memory.

    python tools/prove_big.py [log_n] [kind] [proofs]        # default 24 3 2 -> profiles/r02_prove_2p<log_n>.json
"""
import json
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
import torch
import encrypt_zkvm_b200 as ezk
from tests import _oracle

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
kind = int(sys.argv[2]) if len(sys.argv) > 2 else 3
proofs = int(sys.argv[3]) if len(sys.argv) > 3 else 2
n = 1 << log_n
t0 = time.perf_counter()
prog, ex = ezk.synthetic_case(kind, log_n)
trace, program_hash, outputs = ex.trace(), prog.hash(), ex.outputs()
vm_s = time.perf_counter() - t0
dev = torch.from_numpy(trace.view(np.int64)).to("cuda:0")
torch.cuda.synchronize()
free0, total = torch.cuda.mem_get_info()
res = {"log_n": log_n, "kind": kind, "host_vm_s": round(vm_s, 2), "trace_bytes": int(trace.nbytes)}
with ezk.ExecutionProver(ezk.ProofOptions(), program_hash, outputs, ezk.ServerKey()) as p:
    ms = []
    for _ in range(proofs):
        p.timer_start()
        proof = p.prove_device(dev.data_ptr(), n).to_bytes()
        ms.append(p.timer_stop())
    free1, _ = torch.cuda.mem_get_info()
    res.update({"device_ms": [round(x, 2) for x in ms], "proof_bytes": len(proof),
                "stages_ms": {k: round(v, 2) for k, v in p.stage_times_ms().items()},
                "prover_device_memory_GB": round((free0 - free1) / 1e9, 2), "device_total_GB": round(total / 1e9, 1)})
    p.verify(proof)
    res["product_verifier"] = "accepted"
res["oracle_verifier"] = "accepted" if _oracle.load().verify(proof, program_hash + outputs) == 0 else "REJECTED"
(ROOT / "profiles" / f"r02_prove_2p{log_n}.json").write_text(json.dumps(res, indent=1) + "\n")
print(json.dumps(res), flush=True)
