"""CPU tests of the host side of the product: the C-ABI surface (loads, exports every declared symbol, fails
loudly without a GPU) and the host VM / assembler mirror of the reference's `vm` crate, written after the
reference's own unit tests (vm/src/program/tests, vm/src/processor/tests).  No CUDA compute is called here.
"""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

import encrypt_zkvm_b200 as ezk
from encrypt_zkvm_b200 import _lib, vm
from tests._frames import sponge_hash

ROOT = Path(__file__).resolve().parent.parent
M = 2**128 - 45 * 2**40 + 1


# ------------------------------------------------------------------------------------------------ C ABI
def test_library_exports_every_symbol_declared_in_the_header():
    header = (ROOT / "include" / "ezkvm_prover.h").read_text()
    declared = set(re.findall(r"\b(ezk_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 30
    lib = C.CDLL(str(_lib.LIB_PATH))
    missing = [name for name in sorted(declared) if not hasattr(lib, name)]
    assert not missing, f"declared in include/ezkvm_prover.h but not exported: {missing}"
    # and the Python binding covers the same set
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)


def test_version_and_default_options():
    assert _lib.lib.ezk_version().decode().startswith("encrypt-zkvm-b200")
    opt = _lib.EzkOptions()
    _lib.lib.ezk_default_options(C.byref(opt))
    # ProofOptions::new(32, 8, 0, FieldExtension::None, 8, 127) - vm/src/lib.rs:20
    assert (opt.num_queries, opt.blowup_factor, opt.grinding_factor, opt.field_extension, opt.fri_folding_factor,
            opt.fri_remainder_max_degree) == (32, 8, 0, 1, 8, 127)


def test_host_field_arithmetic_against_big_integers():
    """csrc/field/f128_host.h (transcript, per-proof scalars) and the sponge's x86-64 product and addition chain
    (csrc/host/vm.cc), as compiled into the library: products, squares and x^INV_ALPHA against Python integers."""
    import random
    from encrypt_zkvm_b200.prover import array_to_elements, elements_to_array
    from tests._frames import INV_ALPHA
    rng = random.Random(0xF128)
    edge = [0, 1, 2, M - 1, M - 2, 1 << 64, (1 << 64) - 1, 1 << 127, (45 << 40) - 1, 45 << 40, M - (45 << 40), M >> 1, (M >> 1) + 1]
    a = [x for x in edge for _ in edge] + [rng.randrange(M) for _ in range(3000)]
    b = [y for _ in edge for y in edge] + [rng.randrange(M) for _ in range(3000)]
    a += [M - 1 - rng.randrange(1 << 20) for _ in range(500)] + [rng.randrange(1 << rng.randrange(1, 128)) for _ in range(500)]
    b += [M - 1 - rng.randrange(1 << 20) for _ in range(500)] + [rng.randrange(M) for _ in range(500)]
    ea, eb = elements_to_array(a), elements_to_array(b)
    out = np.empty((4 * len(a), 2), dtype=np.uint64)
    assert _lib.lib.ezk_selftest_host_field(ea.ctypes.data, eb.ctypes.data, len(a), out.ctypes.data) == 0
    got = array_to_elements(out)
    for i, (x, y) in enumerate(zip(a, b)):
        assert got[4 * i] == x * y % M and got[4 * i + 1] == x * y % M, (x, y)
        assert got[4 * i + 2] == x * x % M, x
        assert got[4 * i + 3] == pow(x, INV_ALPHA, M), x


@pytest.mark.parametrize("threads", [1, 2, 4, 7])
def test_threaded_copy_of_the_staged_upload(threads):
    """csrc/host/copy_pool.h inside the built library: part boundaries, unaligned ends, thread start and stop."""
    for size in (1, 4096, 300_000, (8 << 20) + 13):
        assert _lib.lib.ezk_selftest_copy_pool(threads, size) == 0, _lib.lib.ezk_last_error().decode()


def _launch_groups(columns, cap, upload_us, compute_us):
    import ctypes as C
    sizes, n, idle = (C.c_uint32 * 64)(), C.c_uint32(), C.c_uint64()
    rc = _lib.lib.ezk_selftest_launch_groups(columns, cap, upload_us, compute_us, sizes, C.byref(n), C.byref(idle))
    assert rc == 0, _lib.lib.ezk_last_error().decode()
    return list(sizes[: n.value]), idle.value


def test_launch_groups_of_a_host_trace():
    """csrc/host/launch_groups.h (how many columns of a host trace go into one interpolation + LDE launch), replayed by
    the library against a discrete-event model: every column is launched exactly once, no launch exceeds the cap, the
    groups grow while the upload runs ahead of the transforms and stay at one column when the upload is the bottleneck."""
    for cap in (1, 2, 4, 8, 14, 28):
        for up, comp in [(400, 500), (335, 450), (600, 450), (100, 500), (500, 100), (1, 1000), (1000, 1)]:
            sizes, _ = _launch_groups(28, cap, up, comp)
            assert sum(sizes) == 28 and max(sizes) <= cap and min(sizes) >= 1
    sizes, idle = _launch_groups(28, 8, 400, 500)          # compute-bound: growing groups, the GPU never waits
    assert sizes[0] == 1 and idle == 0 and max(sizes) >= 3 and len(sizes) <= 16
    assert all(b >= a for a, b in zip(sizes, sizes[1:-1]))  # non-decreasing (the last launch takes what is left)
    sizes, _ = _launch_groups(28, 8, 100, 500)              # upload far ahead: the cap is reached
    assert max(sizes) == 8 and len(sizes) <= 8
    sizes, _ = _launch_groups(28, 8, 600, 450)              # upload-bound: one column per launch, shortest tail
    assert sizes == [1] * 28
    assert _lib.lib.ezk_selftest_launch_groups(0, 8, 1, 1, None, None, None) != 0


def test_no_cpu_fallback_without_a_device():
    """On a box without a GPU every compute entry point must fail with EZK_ERR_NO_DEVICE, never compute on the CPU."""
    if ezk.device_count() > 0:
        pytest.skip("a CUDA device is visible")
    with pytest.raises(_lib.EzkError) as ei:
        ezk.ExecutionProver(ezk.ProofOptions(), [0, 0], [0] * 16, ezk.ServerKey())
    assert ei.value.code == _lib.EZK_ERR_NO_DEVICE
    trace = np.zeros((28, 64, 2), dtype=np.uint64)
    cols = (C.c_void_p * 28)(*[trace[c].ctypes.data for c in range(28)])
    t = _lib.EzkTrace(C.cast(cols, C.POINTER(C.c_void_p)), 28, 64)
    pi = ezk.PublicInputs([0, 0], [0] * 16, ezk.ServerKey()).to_c()
    opt = ezk.ProofOptions().to_c()
    out, out_len = C.c_void_p(), C.c_size_t()
    rc = _lib.lib.ezk_prove(C.byref(t), C.byref(pi), C.byref(opt), C.byref(out), C.byref(out_len))
    assert rc == _lib.EZK_ERR_NO_DEVICE and not out.value
    assert b"no CUDA device" in _lib.lib.ezk_last_error()


def test_product_package_never_imports_the_oracle():
    for path in (ROOT / "encrypt_zkvm_b200").rglob("*"):
        if path.name == "build.py":  # builds the checker (allowed); never loads it
            continue
        if path.suffix in (".py", ".cu", ".cuh", ".cc", ".h"):
            text = path.read_text()
            assert "liborc" not in text and "oracle/" not in text.replace("the oracle/", ""), path


# ------------------------------------------------------------------------------------------------ assembler
PADDED = "push(1) noop noop noop noop noop noop noop push(2) add read mul noop noop noop noop"


def test_compile_program_display():
    """vm/src/program/tests/mod.rs:70-83 (and :11-60 through Program.load)."""
    assert str(ezk.Program.compile("push.1\npush.2\nadd\nread\nmul")) == PADDED
    assert str(ezk.Program.compile("# Comment 1\npush.1\npush.2 # Comment 2\nadd\nread\nmul\n")) == PADDED


def test_load_program(tmp_path):
    p = tmp_path / "prog.txt"
    p.write_text("push.1\npush.2\nadd\nread\nmul\n")
    assert str(ezk.Program.load(p)) == PADDED
    with pytest.raises(ezk.ProgramError) as ei:
        ezk.Program.load(tmp_path / "missing.txt")
    assert ei.value.message.startswith("program error at 0: ")


def test_program_padding():
    """vm/src/program/tests/mod.rs:85-96."""
    code = ezk.Program.compile("push.1\npush.2\nadd\nread\nread\nread\nmul\nadd\nadd").code()
    assert len(code) % 16 == 0
    assert code[8] == (vm.OPCODES["push"], 2)
    assert code[14] == (vm.OPCODES["noop"], 0) and code[15] == (vm.OPCODES["noop"], 0)


@pytest.mark.parametrize("source,message", [
    ("", "program error at 0: a program must contain at least one instruction"),
    ("push.1\npush.2\nad", "program error at 3: instruction ad is invalid"),
    ("push", "program error at 1: malformed instruction push, parameter is missing"),
    ("push.1.2", "program error at 1: malformed instruction push, too many parameters provided"),
    ("push.abc", "program error at 1: malformed instruction push, parameter 'abc' is invalid"),
    ("add.1", "program error at 1: malformed instruction add, too many parameters provided"),
    ("read\nread2.7", "program error at 2: malformed instruction read2, too many parameters provided"),
    ("noop", "program error at 1: instruction noop is invalid"),  # vm/src/program/mod.rs:107-122: noop is not parseable
])
def test_program_errors(source, message):
    """vm/src/program/errors.rs:11-58, vm/src/program/tests/{mod,parsers}.rs."""
    with pytest.raises(ezk.ProgramError) as ei:
        ezk.Program.compile(source)
    assert ei.value.message == message


def test_program_hash_is_the_rescue_sponge_over_the_padded_code():
    """vm/src/program/mod.rs:88-95 with crypto/src/rescue.rs:42-60 (big-int model in tests/_frames.py)."""
    prog = ezk.Program.compile("push.5\npush.3\nadd")
    assert prog.hash() == sponge_hash(prog.code())


def test_sponge_chain_over_long_programs_and_the_trace_columns_it_feeds():
    """The VM computes x^INV_ALPHA (crypto/src/rescue.rs:146-150,199) with a fixed addition chain on four lanes
    and keeps the per-operation states of Program::compile's hashing pass for the chiplet columns
    (vm/src/processor/chiplets.rs:92-112).  Both must equal the plain big-int sponge, step by step."""
    from tests._frames import apply_round
    for kind in (1, 2, 3):
        prog, ex = ezk.synthetic_case(kind, 12)
        code = prog.code()
        assert len(code) >= 1024 and prog.hash() == sponge_hash(code)
        t = ex.trace()
        state = [0, 0, 0, 0]
        for step, (op, value) in enumerate(code):
            assert _row(t, step)[7:11] == state
            if step % 16 < 14:
                state = apply_round(state, op, value, step)
            else:
                state[2] = state[3] = 0
        assert _row(t, len(code))[7:11] == state and state[:2] == prog.hash()


# ------------------------------------------------------------------------------------------------ processor
def _run(source, public=(), secret=()):
    prog = ezk.Program.compile(source)
    return prog, ezk.execute(prog, ezk.ProgramInputs(list(public), list(secret), ezk.ServerKey()), last_row_seed=5)


def _row(trace, i):
    return [int(trace[c, i, 0]) | (int(trace[c, i, 1]) << 64) for c in range(28)]


def test_trace_layout_row_31():
    """vm/src/processor/tests/mod.rs:19-43."""
    prog, ex = _run("push.5\npush.3\nadd")
    t = ex.trace()
    row = _row(t, 31)
    assert row[0] == 31
    assert row[1:6] == [0] * 5 and row[6] == 0
    assert row[7:9] == prog.hash() and row[9:11] == [0, 0]
    assert row[11] == 1 and row[12] == 8


def test_trace_length_and_random_last_row():
    """vm/src/processor/mod.rs:74,86-92: n = (chiplets capacity + 1).next_power_of_two(); last row randomised."""
    _, ex = _run("push.5\npush.3\nadd")
    t = ex.trace()
    n = t.shape[1]
    assert n == 64 and ex.length == n
    last = _row(t, n - 1)
    assert all(0 < v < M for v in last)
    _, ex2 = _run("push.5\npush.3\nadd")
    assert np.array_equal(ex2.trace(), t)  # seeded stand-in for thread_rng: deterministic


def test_stack_operations():
    """vm/src/processor/tests/stack.rs: mul, add, push, read, noop fill."""
    _, ex = _run("push.2\npush.2\nmul")
    t = ex.trace()
    # PUSH is aligned to 8: ops at 0 (push) and 8 (push), mul at 9
    assert _row(t, 9)[11] == 2 and _row(t, 10)[11] == 1 and _row(t, 10)[12] == 4
    _, ex = _run("read\nread\nadd", public=[7, 9])
    t = ex.trace()
    assert _row(t, 3)[11] == 1 and _row(t, 3)[12] == 16
    assert ex.outputs()[0] == 16


def test_ciphertext_operations_follow_server_key_algebra():
    """fhe/src/server_key.rs:89-124 on the stack (vm/src/processor/stack.rs:155-218)."""
    key = ezk.ServerKey()
    delta = key.parameters.delta
    ct = [11, 22, 33, 44, 55]
    ct2 = [5, 4, 3, 2, 1]
    _, ex = _run("read2\nread\nsmul", public=[3], secret=[ct])
    assert ex.outputs()[:5] == [3 * v % M for v in ct]
    _, ex = _run("read2\nread\nsadd", public=[3], secret=[ct])
    assert ex.outputs()[:5] == ct[:4] + [(ct[4] + delta * 3) % M]
    _, ex = _run("read2\nread2\nadd2", secret=[ct, ct2])
    assert ex.outputs()[:5] == [(a + b) % M for a, b in zip(ct, ct2)]


@pytest.mark.parametrize("source,public,message", [
    ("add", [], "stack error at 1: add operation stack underflow"),
    ("push.1\nmul", [], "stack error at 2: mul operation stack underflow"),
    ("read", [], "stack error at 1: no more inputs to read"),
    ("read2", [], "stack error at 1: no more inputs to read2"),
])
def test_processor_errors(source, public, message):
    """vm/src/processor/errors.rs:14-44 and the *_error tests of vm/src/processor/tests/stack.rs."""
    with pytest.raises(ezk.ProcessorError) as ei:
        _run(source, public=public)
    assert ei.value.message == message


def test_lwe_round_trip_and_homomorphic_ops():
    """fhe/src/tests.rs:6-127 (seeded instead of thread_rng)."""
    params = ezk.LweParameters()
    key = vm.lwe_keygen(params, seed=42)
    assert len(key.key) == 4 and all(v in (0, 1) for v in key.key)
    for m in range(8):
        assert vm.lwe_decrypt(key, vm.lwe_encrypt(key, m, seed=m)) == m
    a, b = vm.lwe_encrypt(key, 2, seed=1), vm.lwe_encrypt(key, 3, seed=2)
    assert vm.lwe_decrypt(key, [(x + y) % M for x, y in zip(a, b)]) == 5
    assert vm.lwe_decrypt(key, [x * 3 % M for x in a]) == 6
    assert vm.lwe_decrypt(key, a[:4] + [(a[4] + params.delta * 4) % M]) == 6


def test_synthetic_programs_have_the_named_lengths_and_valid_traces(oracle):
    """BASELINE.json configs: scalar (1), ciphertext (2), mixed (3) programs padded to 2^k rows."""
    for kind in (1, 2, 3):
        for log_n in (14, 16):  # every row of the larger traces satisfies the AIR too (oracle's row-by-row check)
            prog, ex = ezk.synthetic_case(kind, log_n)
            t = ex.trace()
            assert t.shape == (28, 1 << log_n, 2) and (1 << (log_n - 2)) <= len(prog) < (1 << (log_n - 1))
            assert oracle.validate_trace(t, prog.hash() + ex.outputs()) < 0
    for kind in (1, 2, 3):
        prog, ex = ezk.synthetic_case(kind, 9)
        t = ex.trace()
        assert t.shape == (28, 512, 2)
        assert oracle.validate_trace(t, prog.hash() + ex.outputs()) < 0
        ops = {c for c, _ in prog.code()}
        if kind == 1:
            assert vm.OPCODES["read2"] not in ops and vm.OPCODES["add"] in ops
        if kind == 2:
            assert vm.OPCODES["read2"] in ops and vm.OPCODES["smul"] in ops and vm.OPCODES["add2"] in ops


def test_wire_compat_struct_round_trips_without_a_gpu():
    """ezk_get/set_wire_compat: defaults are the documented reading; the context manager restores them."""
    import ctypes as C
    import encrypt_zkvm_b200 as ezk
    from encrypt_zkvm_b200 import _lib
    cur = _lib.EzkWireCompat()
    _lib.lib.ezk_get_wire_compat(C.byref(cur))
    assert (cur.ood_interleaved, cur.remainder_low_to_high, cur.trace_info_aux_rands_byte, cur.first_nonce) == (1, 1, 1, 1)
    with ezk.wire_compat(ood_interleaved=0, first_nonce=5):
        _lib.lib.ezk_get_wire_compat(C.byref(cur))
        assert (cur.ood_interleaved, cur.remainder_low_to_high, cur.trace_info_aux_rands_byte, cur.first_nonce) == (0, 1, 1, 5)
    _lib.lib.ezk_get_wire_compat(C.byref(cur))
    assert (cur.ood_interleaved, cur.first_nonce) == (1, 1)
    with pytest.raises(ValueError):
        with ezk.wire_compat(no_such_switch=1):
            pass
    _lib.lib.ezk_set_wire_compat(None)


def test_oracle_trace_generator_matches_the_host_vm(oracle):
    """oracle/tracegen.cpp (input of the CPU legs of bench.py, built without the product library) produces the same
    trace, program hash and outputs as ezk_synthetic_case."""
    import encrypt_zkvm_b200 as ezk
    for kind, log_n in [(1, 8), (2, 10), (3, 12)]:
        trace, pub = oracle.synthetic_trace(kind, log_n)
        prog, ex = ezk.synthetic_case(kind, log_n)
        assert np.array_equal(trace, ex.trace())
        assert pub == prog.hash() + ex.outputs()


def test_rust_shim_declares_only_exported_symbols():
    """rust-shim/src/ffi.rs (the reference-side binding of INTEGRATION.md, as files) names only functions that the
    header declares and the library exports, and its constants match the header's status codes."""
    ffi = (ROOT / "rust-shim" / "src" / "ffi.rs").read_text()
    header = (ROOT / "include" / "ezkvm_prover.h").read_text()
    fns = re.findall(r"pub fn (ezk_\w+)\(", ffi)
    assert len(fns) >= 10
    for name in fns:
        assert re.search(rf"\b{name}\(", header), name
        assert hasattr(_lib.lib, name), name
    for name, value in re.findall(r"pub const (EZK_\w+): c_int = (-?\d+);", ffi):
        assert getattr(_lib, name) == int(value), name
