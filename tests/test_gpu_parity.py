"""GPU parity tests: every CUDA stage and the whole proof against the CPU oracle, through the C ABI.

Bar: bit-exact (all arithmetic on this path is integer / byte work).  Sizes are chosen so the oracle finishes
in seconds; larger sizes are covered by size-independent properties in test_gpu_properties.py.
"""
import numpy as np
import pytest

from tests import _oracle
from tests._cases import lr_case, small_case, synthetic, pub_elements

pytestmark = pytest.mark.gpu

M = _oracle.MODULUS


def rand_elems(rng, shape):
    """uniform canonical field elements as uint64 (.., 2) words"""
    lo = rng.integers(0, 2**64, size=shape, dtype=np.uint64)
    hi = rng.integers(0, 2**64 - 1, size=shape, dtype=np.uint64)  # hi < 2^64-1  =>  value < M
    return np.stack([lo, hi], axis=-1)


@pytest.fixture(scope="module")
def prover(gpu_prover_factory):
    ezk = gpu_prover_factory
    p = ezk.ExecutionProver(ezk.ProofOptions(), [0, 0], [0] * 16, ezk.ServerKey())
    yield p
    p.close()


@pytest.mark.parametrize("log_n", [1, 3, 6, 9, 10, 11, 13, 16])
@pytest.mark.parametrize("inverse", [False, True])
def test_ntt_matches_oracle(prover, oracle, log_n, inverse):
    rng = np.random.default_rng(100 + log_n)
    n = 1 << log_n
    cols = rand_elems(rng, (3, n))
    got = prover.stage_ntt(cols, inverse)
    for c in range(3):
        want = oracle.interpolate(cols[c]) if inverse else oracle.forward_ntt(cols[c])
        assert np.array_equal(got[c], want), f"column {c}"


@pytest.mark.parametrize("log_n,width", [(3, 1), (6, 28), (7, 7), (9, 5), (10, 28), (11, 3), (12, 28), (14, 2), (16, 1)])
def test_lde_matches_oracle(prover, oracle, log_n, width):
    rng = np.random.default_rng(200 + log_n)
    n = 1 << log_n
    cols = rand_elems(rng, (width, n))
    got = prover.stage_lde(cols)
    for c in range(width):
        _, want = oracle.lde_column(cols[c])
        assert np.array_equal(got[c], want), f"column {c}"


@pytest.mark.parametrize("order", ["-1", "0", "2", "7"])
def test_lde_is_independent_of_the_cta_order(prover, oracle, monkeypatch, order):
    """The strided NTT passes take their (tile, column) from the linear CTA index (csrc/ntt/ntt.cu, EZK_NTT_ORDER, read
    per launch): column-major (-1), tile-major in groups of 2^k tiles (k beyond log2(tiles) falls back to column-major).
    Every order computes the same table; the default order is what every other test runs."""
    monkeypatch.setenv("EZK_NTT_ORDER", order)
    monkeypatch.setenv("EZK_NTT_FINAL_ORDER", order)
    for log_n, width in [(11, 3), (13, 5), (16, 2)]:  # two-pass plans (one-pass transforms have no strided pass)
        rng = np.random.default_rng(900 + log_n)
        cols = rand_elems(rng, (width, 1 << log_n))
        got = prover.stage_lde(cols)
        for c in range(width):
            _, want = oracle.lde_column(cols[c])
            assert np.array_equal(got[c], want), f"order {order}, 2^{log_n}, column {c}"


@pytest.mark.parametrize("rows,width", [(2, 1), (16, 28), (512, 7), (4096, 8), (1 << 14, 28), (1 << 15, 3)])
def test_merkle_matches_oracle(prover, oracle, rows, width):
    rng = np.random.default_rng(300 + rows)
    table = rand_elems(rng, (width, rows))  # column-major, as the GPU keeps it
    got = prover.stage_merkle(table)
    root, nodes = oracle.merkle_rows(np.ascontiguousarray(table.transpose(1, 0, 2)))
    assert got[32:64] == root
    assert got[32:] == nodes[32:]


@pytest.mark.parametrize("log_s", [4, 10, 13, 16])
def test_fri_fold_matches_formula(prover, oracle, log_s):
    """e'[i] = sum_k (alpha/x_i)^k (1/8) sum_j e[i + j m] zeta^(-jk), x_i = 3 w_s^i."""
    rng = np.random.default_rng(400 + log_s)
    s = 1 << log_s
    m = s // 8
    evals = rand_elems(rng, (s,))
    alpha = int(rng.integers(1, 2**62)) * 0x1234567 % M
    got = _oracle.from_arr(prover.stage_fri_fold(evals, alpha))
    e = _oracle.from_arr(evals)
    w = oracle.root_of_unity(log_s)
    zeta_inv = pow(oracle.root_of_unity(3), M - 2, M)
    inv8, inv3 = pow(8, M - 2, M), pow(3, M - 2, M)
    idx = list(range(0, m, max(1, m // 64)))[:64] + [m - 1]
    for i in idx:
        xinv = inv3 * pow(w, (M - 1 - i) % (M - 1), M) % M
        acc = 0
        for k in range(8):
            sk = sum(e[i + j * m] * pow(zeta_inv, j * k, M) for j in range(8)) % M
            acc = (acc + sk * inv8 % M * pow(alpha * xinv % M, k, M)) % M
        assert got[i] == acc, i


def _check_full_proof(ezk, oracle, trace, program_hash, outputs, delta=16, options=None, min_security=95):
    pub = pub_elements(program_hash, outputs)
    opt = options or ezk.ProofOptions()
    oopt = _oracle.default_options(delta=delta, num_queries=opt.num_queries, grinding=opt.grinding_factor,
                                   fri_rem_max_deg=opt.fri_remainder_max_degree)
    want = oracle.prove(trace, pub, oopt)
    params = ezk.LweParameters(plaintext_modulus=8, ciphertext_modulus=8 * delta)
    with ezk.ExecutionProver(opt, program_hash, outputs, ezk.ServerKey(params)) as p:
        proof = p.prove(trace)
        n = trace.shape[1]
        L = 8 * n
        # stage by stage, in pipeline order, so the first divergence is the one reported
        tl = np.frombuffer(p.artifact("trace_lde"), dtype=np.uint64).reshape(28, L, 2)
        want_tl = want.array("trace_lde").reshape(L, 28, 2).transpose(1, 0, 2)
        assert np.array_equal(tl, want_tl), "trace LDE"
        assert p.artifact("trace_root") == want.raw("trace_root"), "trace root"
        assert p.artifact("combined") == want.raw("combined"), "constraint evaluations"
        cl = np.frombuffer(p.artifact("constraint_lde"), dtype=np.uint64).reshape(7, L, 2)
        want_cl = want.array("comp_lde").reshape(L, 7, 2).transpose(1, 0, 2)
        assert np.array_equal(cl, want_cl), "composition LDE"
        assert p.artifact("constraint_root") == want.raw("comp_root"), "constraint root"
        ood = _oracle.from_arr(np.frombuffer(p.artifact("ood_trace"), dtype=np.uint64))
        assert ood[0::2] == want.elements("ood_cur") and ood[1::2] == want.elements("ood_next"), "OOD trace frame"
        assert p.artifact("ood_constraints") == want.raw("ood_comp"), "OOD constraint evaluations"
        assert p.artifact("deep_evals") == want.raw("deep_evals"), "DEEP evaluations"
        assert p.artifact("fri_roots") == want.raw("fri_roots"), "FRI layer roots"
        assert p.artifact("remainder") == want.raw("remainder"), "FRI remainder"
        assert p.artifact("positions") == want.raw("positions"), "query positions"
    assert proof.to_bytes() == want.proof, "serialized proof bytes"
    assert oracle.verify(proof.to_bytes(), pub, oopt, min_security) == 0, "oracle verifier rejects the GPU proof"
    return proof


def test_prove_linear_regression_example(gpu_prover_factory, oracle):
    """configs[0]: examples/linear_regression (lr.txt, n = 128, 0 FRI layers)."""
    case = lr_case()
    _check_full_proof(gpu_prover_factory, oracle, case.trace, case.program_hash, case.outputs)


def test_prove_reference_unit_test_program(gpu_prover_factory, oracle):
    """vm/src/lib.rs:47-99 `test_prove`: read2 read sadd push.1 push.2 add smul (n = 128 after padding)."""
    case = small_case()
    _check_full_proof(gpu_prover_factory, oracle, case.trace, case.program_hash, case.outputs)


@pytest.mark.parametrize("kind,log_n", [(1, 7), (2, 8), (3, 9), (1, 10), (2, 11), (3, 12), (2, 13), (1, 14)])
def test_prove_synthetic_matches_oracle(gpu_prover_factory, oracle, kind, log_n):
    case = synthetic(kind, log_n)
    assert case.trace.shape[1] == 1 << log_n
    _check_full_proof(gpu_prover_factory, oracle, case.trace, case.program_hash, case.outputs)


def test_prove_other_options_and_delta(gpu_prover_factory, oracle):
    ezk = gpu_prover_factory
    case = synthetic(3, 10, delta=32)
    opt = ezk.ProofOptions(num_queries=20, grinding_factor=8, fri_remainder_max_degree=31)
    # 20 queries * 3 bits - 1 = 59 bits of conjectured security: winter-air counts the grinding bits only once the
    # queries alone give 80 (GRINDING_CONTRIBUTION_FLOOR)
    _check_full_proof(ezk, oracle, case.trace, case.program_hash, case.outputs, delta=32, options=opt, min_security=59)


def test_prove_device_resident_trace_gives_same_bytes(gpu_prover_factory, oracle):
    import torch
    ezk = gpu_prover_factory
    case = synthetic(2, 10)
    with ezk.ExecutionProver(ezk.ProofOptions(), case.program_hash, case.outputs, ezk.ServerKey()) as p:
        host = p.prove(case.trace).to_bytes()
        d = torch.from_numpy(case.trace.view(np.int64)).to("cuda:0")
        torch.cuda.synchronize()
        dev = p.prove_device(d.data_ptr(), case.trace.shape[1]).to_bytes()
        again = p.prove(case.trace).to_bytes()
    assert host == dev == again


def test_invalid_trace_is_rejected(gpu_prover_factory):
    """A trace violating the AIR must fail loudly (winterfell: MismatchedConstraintPolynomialDegree / debug asserts)."""
    ezk = gpu_prover_factory
    case = synthetic(1, 8)
    bad = case.trace.copy()
    bad[12, 5, 0] ^= np.uint64(1)  # flip one stack cell
    with ezk.ExecutionProver(ezk.ProofOptions(), case.program_hash, case.outputs, ezk.ServerKey()) as p:
        with pytest.raises(ezk.ProverError) as ei:
            p.prove(bad)
        assert ei.value.code in (-3, -4)


def test_prover_state_survives_other_trace_lengths(gpu_prover_factory):
    """The prover keeps per-length tables across proofs (periodic columns, boundary divisors, twiddle tables): a call
    with another trace length in between — here one that is rejected — must not leak into the next proof."""
    ezk = gpu_prover_factory
    a, b = synthetic(2, 10), synthetic(1, 11)
    with ezk.ExecutionProver(ezk.ProofOptions(), a.program_hash, a.outputs, ezk.ServerKey()) as p:
        first = p.prove(a.trace).to_bytes()
        with pytest.raises(ezk.ProverError):
            p.prove(b.trace)  # other length, and not a trace of this program
        second = p.prove(a.trace).to_bytes()
        with pytest.raises(ezk.ProverError):
            p.prove(b.trace[:, :512])
        third = p.prove(a.trace).to_bytes()
    assert first == second == third


def test_two_provers_on_one_gpu_from_two_threads(gpu_prover_factory, oracle):
    """Proof-level pipelining: provers share nothing, so two host threads may prove at the same time on
    the same GPU; every proof must still be the oracle's bytes."""
    import threading
    ezk = gpu_prover_factory
    cases = [synthetic(2, 11), synthetic(1, 10)]
    want = [oracle.prove(c.trace, pub_elements(c.program_hash, c.outputs), _oracle.default_options()).proof for c in cases]
    got = [[], []]

    def work(k):
        with ezk.ExecutionProver(ezk.ProofOptions(), cases[k].program_hash, cases[k].outputs, ezk.ServerKey()) as p:
            for _ in range(6):
                got[k].append(p.prove(cases[k].trace).to_bytes())

    threads = [threading.Thread(target=work, args=(k,)) for k in range(2)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for k in range(2):
        assert len(got[k]) == 6 and all(g == want[k] for g in got[k])


def test_argument_errors(gpu_prover_factory):
    ezk = gpu_prover_factory
    case = synthetic(1, 7)
    with ezk.ExecutionProver(ezk.ProofOptions(field_extension=2), case.program_hash, case.outputs, ezk.ServerKey()) as p:
        with pytest.raises(ezk.ProverError) as ei:
            p.prove(case.trace)
        assert ei.value.code == -2  # UnsupportedFieldExtension
    with ezk.ExecutionProver(ezk.ProofOptions(), case.program_hash, case.outputs, ezk.ServerKey()) as p:
        with pytest.raises(ezk.ProverError):
            p.prove(case.trace[:, :96])  # not a power of two


def test_air_frames_on_gpu_match_oracle(prover, oracle):
    """The reference's AIR unit-test frames (air/src/tests/mod.rs) + random frames: 20 values, bit-exact."""
    from tests._frames import reference_frames
    rng = np.random.default_rng(7)
    frames = reference_frames()
    cur = [f.cur for f in frames] + [_oracle.from_arr(rand_elems(rng, (28,))) for _ in range(64)]
    nxt = [f.nxt for f in frames] + [_oracle.from_arr(rand_elems(rng, (28,))) for _ in range(64)]
    per = [f.periodic for f in frames] + [_oracle.from_arr(rand_elems(rng, (9,))) for _ in range(64)]
    for delta in (16, 4096):
        got = prover.stage_eval_frames(np.stack([_oracle.to_arr(c) for c in cur]), np.stack([_oracle.to_arr(c) for c in nxt]),
                                       np.stack([_oracle.to_arr(c) for c in per]), delta)
        for k in range(len(cur)):
            want = oracle.evaluate_transition(cur[k], nxt[k], per[k], delta=delta)
            assert _oracle.from_arr(got[k]) == want, f"frame {k}"


def test_air_frames_through_the_production_grouped_path(prover, oracle):
    """The constraint kernel's own path (selector-grouped accumulation, flagged arithmetic + exact redo) at frame
    level: sum_j c_j r_j per frame against the same sum over the oracle's 20 values, for the reference's unit-test
    frames and random frames, so that an AIR regression is localised before the `combined` column."""
    from tests._frames import reference_frames
    rng = np.random.default_rng(11)
    frames = reference_frames()
    cur = [f.cur for f in frames] + [_oracle.from_arr(rand_elems(rng, (28,))) for _ in range(192)]
    nxt = [f.nxt for f in frames] + [_oracle.from_arr(rand_elems(rng, (28,))) for _ in range(192)]
    per = [f.periodic for f in frames] + [_oracle.from_arr(rand_elems(rng, (9,))) for _ in range(192)]
    # random frames with valid op-bit / flag columns as well: selectors 0/1 make single groups survive
    for k in range(len(frames), len(cur), 2):
        for c in range(1, 7):
            cur[k][c] = int(rng.integers(0, 2))
        per[k][0] = int(rng.integers(0, 2))
    tcoef = _oracle.from_arr(rand_elems(rng, (20,)))
    for delta in (16, 4096):
        got = _oracle.from_arr(prover.stage_eval_frames_sum(np.stack([_oracle.to_arr(c) for c in cur]),
                                                            np.stack([_oracle.to_arr(c) for c in nxt]),
                                                            np.stack([_oracle.to_arr(c) for c in per]), delta, tcoef))
        for k in range(len(cur)):
            r = oracle.evaluate_transition(cur[k], nxt[k], per[k], delta=delta)
            want = sum(c * v for c, v in zip(tcoef, r)) % M
            assert got[k] == want, f"frame {k}"
        # one coefficient at a time isolates every constraint inside its group
        for j in range(20):
            unit = [1 if i == j else 0 for i in range(20)]
            one = _oracle.from_arr(prover.stage_eval_frames_sum(np.stack([_oracle.to_arr(c) for c in cur[:40]]),
                                                                np.stack([_oracle.to_arr(c) for c in nxt[:40]]),
                                                                np.stack([_oracle.to_arr(c) for c in per[:40]]), delta, unit))
            for k in range(40):
                assert one[k] == oracle.evaluate_transition(cur[k], nxt[k], per[k], delta=delta)[j], (j, k)


@pytest.mark.parametrize("flags", [dict(compat_ood_interleaved=0), dict(compat_remainder_low_to_high=0),
                                   dict(compat_trace_info_aux_rands_byte=0), dict(compat_first_nonce=0),
                                   dict(compat_ood_interleaved=0, compat_remainder_low_to_high=0,
                                        compat_trace_info_aux_rands_byte=0, compat_first_nonce=7)])
def test_wire_compat_switches_follow_the_oracle(gpu_prover_factory, oracle, flags):
    """The [V] items of SURVEY App. A.13 sit behind ONE struct in the product (ezk_wire_compat) that mirrors the
    oracle's compat_* flags: flipped together, writer and verifier of both sides still agree byte for byte; a proof
    written under one reading is not accepted under another when the reading reaches a hashed or parsed value."""
    ezk = gpu_prover_factory
    case = synthetic(2, 10)
    pub = pub_elements(case.program_hash, case.outputs)
    oopt = _oracle.default_options(**flags)
    want = oracle.prove(case.trace, pub, oopt).proof
    product = {k[len("compat_"):]: v for k, v in flags.items()}
    with ezk.ExecutionProver(ezk.ProofOptions(), case.program_hash, case.outputs, ezk.ServerKey()) as p:
        default_bytes = p.prove(case.trace).to_bytes()
        with ezk.wire_compat(**product):
            got = p.prove(case.trace).to_bytes()
            p.verify(got)
            assert oracle.verify(got, pub, oopt) == 0
        assert got == want
        assert got != default_bytes
        assert p.prove(case.trace).to_bytes() == default_bytes  # the switches are restored
        if set(flags) != {"compat_first_nonce"}:  # any valid nonce verifies: only the search start differs
            with pytest.raises(ezk.VerifierError):
                p.verify(got)


def test_non_canonical_trace_elements_are_rejected(gpu_prover_factory):
    """`BaseElement` memory is always canonical (< M); bytes that are not must fail loudly, not prove garbage."""
    ezk = gpu_prover_factory
    case = synthetic(1, 8)
    bad = case.trace.copy()
    bad[3, 17] = (np.uint64(0xFFFFD30000000001), np.uint64(0xFFFFFFFFFFFFFFFF))  # = M
    with ezk.ExecutionProver(ezk.ProofOptions(), case.program_hash, case.outputs, ezk.ServerKey()) as p:
        with pytest.raises(ezk.ProverError) as ei:
            p.prove(bad)
        assert ei.value.code == -1 and "non-canonical" in ei.value.message
        assert p.prove(case.trace).to_bytes()  # the prover stays usable


def test_trace_shape_errors(gpu_prover_factory):
    """Edge shapes: too short, wrong width (the AIR fixes 28 columns), zero rows."""
    import ctypes as C
    from encrypt_zkvm_b200 import _lib
    ezk = gpu_prover_factory
    case = synthetic(1, 7)
    with ezk.ExecutionProver(ezk.ProofOptions(), case.program_hash, case.outputs, ezk.ServerKey()) as p:
        with pytest.raises(ezk.ProverError):
            p.prove(case.trace[:, :32])  # n = 32 < 64
        with pytest.raises(ezk.ProverError):
            p.prove(case.trace[:, :0])
        with pytest.raises(ValueError):
            p.prove(case.trace[:27])
        # wrong width through the raw C ABI
        a = np.ascontiguousarray(case.trace[:27])
        cols = (C.c_void_p * 27)(*[a[c].ctypes.data for c in range(27)])
        t = _lib.EzkTrace(C.cast(cols, C.POINTER(C.c_void_p)), 27, a.shape[1])
        out, out_len = C.c_void_p(), C.c_size_t()
        rc = _lib.lib.ezk_prover_prove(p._handle, C.byref(t), C.byref(p.pub_inputs.to_c()), C.byref(p.options.to_c()),
                                       C.byref(out), C.byref(out_len))
        assert rc == _lib.EZK_ERR_INVALID_ARGUMENT


@pytest.mark.parametrize("log_n", [6, 10, 13])
def test_transforms_on_values_near_the_modulus(prover, oracle, log_n):
    """Elements whose top limb is all ones (M - 1 - k: small negatives) take the kernels' rare-tail path
    (flagged arithmetic -> exact recomputation of the butterfly group); results must still be bit-exact."""
    rng = np.random.default_rng(900 + log_n)
    n = 1 << log_n
    vals = [M - 1 - int(k) for k in rng.integers(0, 1 << 40, size=n)]
    vals[0], vals[1], vals[2] = M - 1, 0, 1
    col = _oracle.to_arr(vals).reshape(1, n, 2)
    for inverse in (False, True):
        got = prover.stage_ntt(col, inverse)
        want = oracle.interpolate(col[0]) if inverse else oracle.forward_ntt(col[0])
        assert np.array_equal(got[0], want), inverse
    _, want_lde = oracle.lde_column(col[0])
    assert np.array_equal(prover.stage_lde(col)[0], want_lde)


@pytest.mark.parametrize("case_fn", [lr_case, small_case, lambda: synthetic(3, 10), lambda: synthetic(2, 13)])
def test_product_verifier_accepts_and_agrees_with_the_oracle(gpu_prover_factory, oracle, case_fn):
    """`ezk_prover_verify` (winterfell::verify for this AIR, vm/src/lib.rs:91-98): accepts the prover's and the
    oracle's proofs, and returns the same accept / reject decision as the oracle verifier on every mutation."""
    ezk = gpu_prover_factory
    case = case_fn()
    pub = pub_elements(case.program_hash, case.outputs)
    with ezk.ExecutionProver(ezk.ProofOptions(), case.program_hash, case.outputs, ezk.ServerKey()) as p:
        proof = p.prove(case.trace).to_bytes()
        p.verify(proof)
        p.verify(oracle.prove(case.trace, pub).proof)
        rng = np.random.default_rng(len(proof))
        offsets = [0, 2, 5, 22, 25, 28, 30, 40, len(proof) - 1, len(proof) - 9] + [int(v) for v in rng.integers(0, len(proof), size=60)]
        for off in offsets:
            bad = bytearray(proof)
            bad[off] ^= 1 << int(rng.integers(0, 8))
            want_ok = oracle.verify(bytes(bad), pub) == 0
            try:
                p.verify(bytes(bad))
                got_ok = True
            except ezk.VerifierError:
                got_ok = False
            assert got_ok == want_ok, off
            assert not got_ok or bytes(bad) == proof
        with pytest.raises(ezk.VerifierError):
            p.verify(proof[:-1])
        with pytest.raises(ezk.VerifierError):
            p.verify(proof + b"\x00")
        with pytest.raises(ezk.VerifierError) as ei:
            p.verify(proof, min_conjectured_security=96)  # 32 queries * 3 bits - 1 = 95
        assert "security" in ei.value.message
    wrong_out = list(case.outputs)
    wrong_out[0] = (wrong_out[0] + 1) % M
    with ezk.ExecutionProver(ezk.ProofOptions(), case.program_hash, wrong_out, ezk.ServerKey()) as p:
        with pytest.raises(ezk.VerifierError):
            p.verify(proof)
    wrong_hash = [(case.program_hash[0] + 1) % M, case.program_hash[1]]
    with ezk.ExecutionProver(ezk.ProofOptions(), wrong_hash, case.outputs, ezk.ServerKey()) as p:
        with pytest.raises(ezk.VerifierError):
            p.verify(proof)


def test_product_verifier_on_a_2p18_proof_and_other_options(gpu_prover_factory):
    ezk = gpu_prover_factory
    case = synthetic(2, 18)
    with ezk.ExecutionProver(ezk.ProofOptions(), case.program_hash, case.outputs, ezk.ServerKey()) as p:
        p.verify(p.prove(case.trace))
    case = synthetic(3, 10, delta=32)
    opt = ezk.ProofOptions(num_queries=20, grinding_factor=8, fri_remainder_max_degree=31)
    params = ezk.LweParameters(plaintext_modulus=8, ciphertext_modulus=8 * 32)
    with ezk.ExecutionProver(opt, case.program_hash, case.outputs, ezk.ServerKey(params)) as p:
        proof = p.prove(case.trace)
        p.verify(proof, min_conjectured_security=59)
        with pytest.raises(ezk.VerifierError):
            p.verify(proof, min_conjectured_security=60)  # the 8 grinding bits do not count below 80 query bits
        with pytest.raises(ezk.VerifierError):
            p.verify(proof)  # 59 bits < 95
    with ezk.ExecutionProver(opt, case.program_hash, case.outputs, ezk.ServerKey()) as p:  # delta 16: wrong AIR parameter
        with pytest.raises(ezk.VerifierError):
            p.verify(proof, min_conjectured_security=59)


@pytest.mark.parametrize("which", ["lr", "small", (1, 10), (2, 12), (3, 14)])
def test_bookkeeping_columns_generated_on_the_device(gpu_prover_factory, which):
    """clk, op bits, chiplet flag and stack depth come from the operation list on the device;
    the proof must equal the one from the host VM's full trace, byte for byte, and the eight host columns are never read
    (they are poisoned here).  Reference layout of those columns: vm/src/processor/{system,decoder,chiplets,stack}.rs."""
    ezk = gpu_prover_factory
    case = lr_case() if which == "lr" else small_case() if which == "small" else synthetic(*which)
    codes = case.program.op_codes()
    n = case.trace.shape[1]
    assert len(codes) < n
    poisoned = case.trace.copy()
    for c in ezk.ExecutionProver.OP_COLUMNS:
        poisoned[c] = 0xFFFFFFFFFFFFFFFF
    with ezk.ExecutionProver(ezk.ProofOptions(), case.program_hash, case.outputs, case.key) as p:
        want = p.prove(case.trace).to_bytes()
        got = p.prove_with_ops(poisoned, codes, last_row=case.trace[:, n - 1]).to_bytes()
        assert got == want
        with pytest.raises(ezk.ProverError):  # an unknown opcode is rejected, not proved
            bad = codes.copy()
            bad[3] = 0b11111
            p.prove_with_ops(poisoned, bad, last_row=case.trace[:, n - 1])
        with pytest.raises(ezk.ProverError):  # a wrong operation makes the trace violate the AIR
            bad = codes.copy()
            bad[0] = 0b10000 if codes[0] != 0b10000 else 0b10001
            p.prove_with_ops(poisoned, bad, last_row=case.trace[:, n - 1])
        assert p.prove_with_ops(poisoned, codes, last_row=case.trace[:, n - 1]).to_bytes() == want
