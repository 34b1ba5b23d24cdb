"""Torch-free quick check: host-trace proofs at 2^log_n rows through ExecutionProver.prove (wall clock per proof,
host<->device copies included) and the device-side time of the same calls.

    python tools/e2e_quick.py [log_n] [kind] [proofs]
"""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import encrypt_zkvm_b200 as ezk

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
kind = int(sys.argv[2]) if len(sys.argv) > 2 else 2
proofs = int(sys.argv[3]) if len(sys.argv) > 3 else 5
t0 = time.perf_counter()
prog, ex = ezk.synthetic_case(kind, log_n)
trace, program_hash, outputs = ex.trace(), prog.hash(), ex.outputs()
t_vm = time.perf_counter() - t0
with ezk.ExecutionProver(ezk.ProofOptions(), program_hash, outputs, ezk.ServerKey()) as p:
    for _ in range(3):
        proof = p.prove(trace)
    wall, dev = [], []
    for _ in range(proofs):
        p.timer_start()
        t0 = time.perf_counter()
        proof = p.prove(trace)
        wall.append((time.perf_counter() - t0) * 1e3)
        dev.append(p.timer_stop())
    stages = p.stage_times_ms()
    p.verify(proof)
print(f"log_n={log_n} vm_s={t_vm:.2f} proof_bytes={len(proof.to_bytes())} e2e_ms={sorted(wall)[len(wall) // 2]:.2f} "
      f"(min {min(wall):.2f}) device_ms={sorted(dev)[len(dev) // 2]:.2f} verified=1")
print("stages:", {k: round(v, 3) for k, v in stages.items()})
