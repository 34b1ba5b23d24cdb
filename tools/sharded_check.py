"""Run under torchrun (one rank per GPU): proves the same trace on every rank alone, then as ONE proof sharded over
all ranks (ExecutionProver.join_group), checks the bytes are identical and prints both device times.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 \
        tools/sharded_check.py [log_n ...]
"""
import hashlib
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import torch.distributed as dist
import encrypt_zkvm_b200 as ezk

rank, local_rank, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local_rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
logs = [int(x) for x in sys.argv[1:]] or [8, 12, 16]
ok = True
for log_n in logs:
    prog, ex = ezk.synthetic_case(2, log_n)
    trace = ex.trace()
    import numpy as np
    dev = torch.from_numpy(trace.view(np.int64)).to(f"cuda:{local_rank}")  # resident copy: device time without PCIe
    torch.cuda.synchronize()
    n = trace.shape[1]
    with ezk.ExecutionProver(ezk.ProofOptions(), prog.hash(), ex.outputs(), ezk.ServerKey(), device=local_rank) as p:
        single = p.prove(trace).to_bytes()
        assert p.prove_device(dev.data_ptr(), n).to_bytes() == single
        p.timer_start()
        p.prove_device(dev.data_ptr(), n)
        ms_single = p.timer_stop()
        p.join_group()
        sharded = p.prove(trace).to_bytes()
        dist.barrier()
        p.prove_device(dev.data_ptr(), n)
        dist.barrier()
        p.timer_start()
        for _ in range(3):
            sharded_dev = p.prove_device(dev.data_ptr(), n).to_bytes()
        ms_sharded = p.timer_stop() / 3
        sharded = sharded if sharded_dev == sharded else b""
        stages = p.stage_times_ms()
        p.leave_group()
        again = p.prove(trace).to_bytes()
    same = sharded == single and again == single
    ok &= same
    t = torch.tensor([ms_sharded], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"log_n": log_n, "world": world, "identical": same, "proof_sha256": hashlib.sha256(single).hexdigest()[:16],
                          "ms_single_gpu": round(ms_single, 3), "ms_sharded_max_over_ranks": round(float(t.item()), 3),
                          "speedup": round(ms_single / float(t.item()), 2),
                          "stages_rank0": {k: round(v, 3) for k, v in stages.items()}}), flush=True)
flag = torch.tensor([0 if ok else 1], device="cuda")
dist.all_reduce(flag)
dist.destroy_process_group()
sys.exit(1 if int(flag.item()) else 0)
