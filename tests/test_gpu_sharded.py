"""The coset-sharded single proof on ONE GPU: `world` provers of this process, one host thread each, joined
through the in-process group of the C ABI (ezk_local_group_*, host-synchronised device copies in place of NCCL).
Everything the multi-GPU path does - column ownership of the interpolation, coset ownership of the LDE rows, the
all-to-all of leaf digests into per-rank Merkle subtrees with only the subtree roots gathered, per-coset interpolation
of the constraint evaluations + 8-point inverse DFT across cosets, column-sharded out-of-domain frame, row-sharded DEEP
composition and FRI layers, owner-routed query openings - runs here and must give the bytes of the single-GPU proof
(which test_gpu_parity.py compares with the oracle).  The NCCL transport itself is covered by the 2-GPU test below and
by bench.py --gpus N (which checks byte identity with the single-GPU proof on every run)."""
import numpy as np
import pytest

from tests import _oracle
from tests._cases import lr_case, small_case, synthetic

pytestmark = pytest.mark.gpu


def _single(ezk, case, options=None):
    with ezk.ExecutionProver(options or ezk.ProofOptions(), case.program_hash, case.outputs, case.key) as p:
        return p.prove(case.trace).to_bytes()


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("kind,log_n", [(2, 7), (1, 9), (3, 11), (2, 13)])
def test_in_process_group_gives_the_single_gpu_bytes(gpu_prover_factory, oracle, monkeypatch, world, kind, log_n):
    ezk = gpu_prover_factory
    case = synthetic(kind, log_n)
    want = _single(ezk, case)
    assert want == oracle.prove(case.trace, case.program_hash + case.outputs).proof
    # small FRI layers stay sharded too (the default only shards layers of >= 4096 rows)
    monkeypatch.setenv("EZK_SHARD_FRI_MIN_ROWS", "64")
    got = ezk.prove_sharded_in_process(world, ezk.ProofOptions(), case.program_hash, case.outputs, case.key, case.trace)
    assert all(g == want for g in got), [g == want for g in got]


@pytest.mark.parametrize("world", [2, 8])
def test_default_fri_threshold_and_reference_programs(gpu_prover_factory, world):
    ezk = gpu_prover_factory
    for case in (lr_case(), small_case(), synthetic(2, 16)):
        want = _single(ezk, case)
        got = ezk.prove_sharded_in_process(world, ezk.ProofOptions(), case.program_hash, case.outputs, case.key, case.trace)
        assert all(g == want for g in got)


def test_sharded_code_path_with_one_member(gpu_prover_factory, monkeypatch):
    """EZK_FORCE_SHARDED_PATH=1: the sharded pipeline with G = 1 (every exchange is a local copy)."""
    ezk = gpu_prover_factory
    case = synthetic(3, 12)
    want = _single(ezk, case)
    monkeypatch.setenv("EZK_FORCE_SHARDED_PATH", "1")
    monkeypatch.setenv("EZK_SHARD_FRI_MIN_ROWS", "64")
    assert _single(ezk, case) == want


def test_a_rank_reads_only_the_columns_it_owns(gpu_prover_factory):
    """Rank r is handed a trace whose other columns are poisoned with non-canonical bytes; device-resident traces too."""
    import torch
    ezk = gpu_prover_factory
    case = synthetic(2, 12)
    want = _single(ezk, case)
    got = ezk.prove_sharded_in_process(4, ezk.ProofOptions(), case.program_hash, case.outputs, case.key, case.trace,
                                       own_columns_only=True)
    assert all(g == want for g in got)
    # device-resident traces take their own schedule (own columns extended first, foreign columns as they arrive)
    dev = torch.from_numpy(case.trace.view(np.int64)).cuda()
    for world in (2, 4, 8):
        got = ezk.prove_sharded_in_process(world, ezk.ProofOptions(), case.program_hash, case.outputs, case.key, case.trace,
                                           device_ptr=dev.data_ptr())
        assert all(g == want for g in got), world


def test_other_options_and_verifier(gpu_prover_factory, oracle):
    ezk = gpu_prover_factory
    case = synthetic(2, 10)
    opt = ezk.ProofOptions(num_queries=20, grinding_factor=8, fri_remainder_max_degree=31)
    want = _single(ezk, case, opt)
    got = ezk.prove_sharded_in_process(4, opt, case.program_hash, case.outputs, case.key, case.trace)
    assert all(g == want for g in got)
    oopt = _oracle.default_options(num_queries=20, grinding=8, fri_rem_max_deg=31)
    assert oracle.verify(got[0], case.program_hash + case.outputs, oopt, 0) == 0


def test_invalid_trace_fails_on_every_member_instead_of_hanging(gpu_prover_factory):
    """A trace that does not satisfy the AIR (or holds non-canonical bytes in ONE rank's column) is rejected by every
    member: the decision is taken on the gathered flags, so no member is left waiting in a collective."""
    ezk = gpu_prover_factory
    case = synthetic(2, 10)
    bad = case.trace.copy()
    bad[5, 100, 0] ^= np.uint64(1)  # column 5 belongs to rank 1 of 4
    with pytest.raises(ezk.ProverError) as ei:
        ezk.prove_sharded_in_process(4, ezk.ProofOptions(), case.program_hash, case.outputs, case.key, bad)
    assert ei.value.code == -3
    bad = case.trace.copy()
    bad[6, 17] = (np.uint64(0xFFFFD30000000001), np.uint64(0xFFFFFFFFFFFFFFFF))  # = M, in rank 2's column
    with pytest.raises(ezk.ProverError) as ei:
        ezk.prove_sharded_in_process(4, ezk.ProofOptions(), case.program_hash, case.outputs, case.key, bad)
    assert ei.value.code == -1


def test_sharded_proof_is_byte_identical_on_two_gpus_over_nccl(gpu_prover_factory):
    """The same pipeline over NCCL (one process per GPU).  Needs a box with >= 2 GPUs; skipped otherwise."""
    import subprocess
    import sys
    from pathlib import Path
    if gpu_prover_factory.device_count() < 2:
        pytest.skip("needs two GPUs (the in-process tests above cover the pipeline on one)")
    root = Path(__file__).resolve().parent.parent
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29577", str(root / "tools" / "sharded_check.py"),
                        "7", "10", "13"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count('"identical": true') == 3
