"""A/B of the run-time knobs of the NTT passes (environment variables the library reads per launch: EZK_NTT_ORDER,
EZK_NTT_FINAL_ORDER, EZK_NTT_STAGE) inside ONE process: for every setting one warm-up proof and `reps` measured proofs
of a device-resident trace; prints the per-kernel times of the NTT passes, the stage times and the proof digest (every
setting must give the same bytes).

    python tools/ntt_order_ab.py [log_n] [kind] [reps] [NAME=V,NAME=V ...]

EZK_NTT_ORDER: -1 = column-major (plain grid order), k >= 0 = tile-major in groups of 2^k adjacent tiles (see
ntt_strided_pass).  EZK_NTT_STAGE: bit 0 strided / bit 1 final passes stage first-step inputs with cp.async.
"""
import hashlib
import json
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
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
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

settings = ["EZK_NTT_ORDER=-1", "EZK_NTT_ORDER=0", "EZK_NTT_ORDER=1", "EZK_NTT_ORDER=2", "EZK_NTT_ORDER=3", "EZK_NTT_ORDER=5",
            "EZK_NTT_FINAL_ORDER=0", "EZK_NTT_FINAL_ORDER=1", "EZK_NTT_FINAL_ORDER=3", "EZK_NTT_ORDER=-1"]
if len(sys.argv) > 4:
    settings = sys.argv[4:]  # each "NAME=V,NAME=V" (an empty string or "-" = the defaults)
digests = set()
touched = set()
for setting in settings:
    for name in touched:
        os.environ.pop(name, None)
    for kv in setting.split(","):
        if "=" in kv:
            name, v = kv.split("=")
            os.environ[name] = v
            touched.add(name)
    p.prove_device(dev.data_ptr(), 1 << log_n)  # warm-up (tables, caches)
    ezk.profile_enable(True)
    ezk.profile_reset()
    p.timer_start()
    for _ in range(reps):
        proof = p.prove_device(dev.data_ptr(), 1 << log_n)
    ms = p.timer_stop() / reps
    prof = ezk.profile_read()
    ezk.profile_enable(False)
    d = hashlib.sha256(proof.to_bytes()).hexdigest()[:16]
    digests.add(d)
    k = {name: round(v["ms"] / reps, 3) for name, v in prof.items() if name.startswith("ntt_")}
    st = {name: round(v, 3) for name, v in p.stage_times_ms().items() if name in ("trace_lde", "composition", "deep")}
    print(json.dumps({"log_n": log_n, "env": setting, "device_ms": round(ms, 3), "kernels": k,
                      "stages_last": st, "sha256_16": d}), flush=True)
print("identical bytes under every setting:", len(digests) == 1)
p.close()
