"""world_size-2 `gloo` test of the N>1 path's host logic (encrypt_zkvm_b200/parallel.py) on CPU: unit assignment,
barrier-bracketed timing with MAX over ranks, and the gather of proof digests to rank 0.  The GPU prover is
replaced by a stand-in (sha256 of the host-built trace) because no CUDA device exists here; the traces themselves
come from the product's host VM, so distinct units really are distinct workloads."""
import hashlib
import os
import socket
import time

import pytest


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, num_units: int, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import encrypt_zkvm_b200 as ezk
        from encrypt_zkvm_b200 import parallel

        units = parallel.assign_units(num_units, world, rank)

        def prove_one(unit: int) -> bytes:
            _, ex = ezk.synthetic_case(2, 8, seed=parallel.unit_seed(0xE2C0DE00, 8, unit))
            if rank == 1:
                time.sleep(0.05)  # make rank 1 the slow one: the reported time must be ITS time
            return hashlib.sha256(ex.trace().tobytes()).digest()

        rep = parallel.prove_units(units, prove_one, time.perf_counter)
        out.put((rank, units, rep.units, rep.seconds, rep.total_units, rep.throughput))
    finally:
        dist.destroy_process_group()


def test_assign_units_covers_every_unit_once():
    from encrypt_zkvm_b200 import parallel
    for world in (1, 2, 3, 8):
        seen = []
        for r in range(world):
            seen += parallel.assign_units(13, world, r)
        assert sorted(seen) == list(range(13))
    with pytest.raises(ValueError):
        parallel.assign_units(4, 2, 2)
    assert parallel.unit_seed(100, 20, 0) != parallel.unit_seed(100, 20, 1)


def test_two_ranks_over_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port, world, num_units = _free_port(), 2, 5
    procs = [ctx.Process(target=_worker, args=(r, world, port, num_units, out)) for r in range(world)]
    for p in procs:
        p.start()
    results = [out.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    by_rank = {r[0]: r for r in results}
    assert by_rank[0][1] == [0, 2, 4] and by_rank[1][1] == [1, 3]
    merged = by_rank[0][2]
    assert sorted(merged) == list(range(num_units))           # rank 0 holds every unit's digest exactly once
    assert len(set(merged.values())) == num_units             # distinct seeds -> distinct traces
    assert by_rank[0][4] == by_rank[1][4] == num_units        # SUM over ranks
    assert by_rank[0][3] == pytest.approx(by_rank[1][3])      # MAX over ranks, same on both
    assert by_rank[0][3] >= 0.1                               # at least rank 1's two sleeps
    assert by_rank[0][5] == pytest.approx(num_units / by_rank[0][3])


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_shard_layout_of_the_split_merkle_commitment(world):
    """Index arithmetic of the sharded proof on the CPU (csrc/dist/shard_layout.h, shared with the prover's opening
    planner): leaf-digest all-to-all, per-rank subtrees, host-side top levels, node lookup and authentication paths
    agree with the unsplit tree for every supported world size."""
    from encrypt_zkvm_b200 import _lib
    for log_leaves in (7, 10, 13):
        assert _lib.lib.ezk_selftest_shard_layout(world, log_leaves, 1234 + log_leaves) == 0, _lib.lib.ezk_last_error()
    assert _lib.lib.ezk_selftest_shard_layout(3, 10, 1) != 0  # not a supported world size
