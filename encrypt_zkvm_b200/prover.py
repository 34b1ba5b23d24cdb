"""Host-side mirror of the reference's `prover` crate over the C ABI.

Reference interface (prover/src/lib.rs:17-77):
    ExecutionProver::new(options, program_hash[2], stack_outputs[16], &server_key)
    prover.prove(trace: TraceTable<BaseElement>) -> Result<Proof, ProverError>
Here `ExecutionProver(options, program_hash, stack_outputs, server_key).prove(trace)` returns a `Proof`
whose `.to_bytes()` is the serialized `winterfell::Proof`.  Field elements are Python ints in [0, M);
traces are `numpy.uint64` arrays of shape (28, n, 2) = little-endian (lo, hi) words of each element.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import EzkError, EzkOptions, EzkPublicInputs, EzkTrace, check, lib

MODULUS = 2**128 - 45 * 2**40 + 1
TRACE_WIDTH = 28


def elements_to_array(values: Sequence[int]) -> np.ndarray:
    """ints -> (len, 2) uint64 array (16 little-endian bytes per element)."""
    out = np.empty((len(values), 2), dtype=np.uint64)
    for i, v in enumerate(values):
        v = int(v)
        if not 0 <= v < MODULUS:
            raise ValueError("field element out of range")
        out[i, 0] = v & 0xFFFFFFFFFFFFFFFF
        out[i, 1] = v >> 64
    return out


def array_to_elements(a: np.ndarray) -> List[int]:
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 2)
    return [int(lo) | (int(hi) << 64) for lo, hi in a]


def bytes_to_elements(b: bytes) -> List[int]:
    return [int.from_bytes(b[i:i + 16], "little") for i in range(0, len(b), 16)]


@dataclass
class ProofOptions:
    """winterfell::ProofOptions::new(32, 8, 0, FieldExtension::None, 8, 127) - vm/src/lib.rs:20."""
    num_queries: int = 32
    blowup_factor: int = 8
    grinding_factor: int = 0
    field_extension: int = 1  # FieldExtension::None
    fri_folding_factor: int = 8
    fri_remainder_max_degree: int = 127

    def to_c(self) -> EzkOptions:
        return EzkOptions(self.num_queries, self.blowup_factor, self.grinding_factor, self.field_extension,
                          self.fri_folding_factor, self.fri_remainder_max_degree)


@dataclass
class LweParameters:
    """fhe/src/parameters.rs:4-21."""
    plaintext_modulus: int = 8
    ciphertext_modulus: int = 128
    k: int = 4
    std: float = 2.412390240121573e-5

    @property
    def delta(self) -> int:
        return self.ciphertext_modulus // self.plaintext_modulus


@dataclass
class ServerKey:
    """fhe/src/server_key.rs:13-17 (only `parameters` is read by the prover/AIR)."""
    parameters: LweParameters = field(default_factory=LweParameters)
    key: Optional[List[int]] = None

    def lwe_size(self) -> int:
        return self.parameters.k + 1


@dataclass
class PublicInputs:
    """air::PublicInputs::new(program_hash, stack_outputs, server_key) - air/src/lib.rs:18-36."""
    program_hash: Sequence[int]
    stack_outputs: Sequence[int]
    server_key: ServerKey

    def to_elements(self) -> List[int]:  # air/src/lib.rs:38-47
        return list(self.program_hash) + list(self.stack_outputs)

    def to_c(self) -> EzkPublicInputs:
        if len(self.program_hash) != 2 or len(self.stack_outputs) != 16:
            raise ValueError("program_hash must have 2 and stack_outputs 16 elements")
        pi = EzkPublicInputs()
        for i, v in enumerate(self.program_hash):
            pi.program_hash[i][:] = list(int(v).to_bytes(16, "little"))
        for i, v in enumerate(self.stack_outputs):
            pi.stack_outputs[i][:] = list(int(v).to_bytes(16, "little"))
        pi.lwe_k = self.server_key.parameters.k
        pi.lwe_delta = self.server_key.parameters.delta
        return pi


class Proof:
    """Serialized `winterfell::Proof` (`Proof::to_bytes` layout)."""

    def __init__(self, data: bytes):
        self._data = bytes(data)

    def to_bytes(self) -> bytes:
        return self._data

    def __len__(self) -> int:
        return len(self._data)


class VerifierError(EzkError):
    """Mirrors winterfell::VerifierError: the proof was rejected."""


class ProverError(EzkError):
    """Mirrors winterfell::ProverError (the reference unwraps it at vm/src/lib.rs:26)."""


def _as_trace_array(trace) -> np.ndarray:
    a = np.ascontiguousarray(trace, dtype=np.uint64)
    if a.ndim != 3 or a.shape[0] != TRACE_WIDTH or a.shape[2] != 2:
        raise ValueError("trace must have shape (28, n, 2) of uint64 words")
    return a


class ExecutionProver:
    """prover::ExecutionProver over a `GpuProver` (one CUDA device, reusable workspace)."""

    def __init__(self, options: ProofOptions, program_hash: Sequence[int], stack_outputs: Sequence[int],
                 server_key: ServerKey, device: int = 0):
        self.options = options
        self.pub_inputs = PublicInputs(program_hash, stack_outputs, server_key)
        self._handle = C.c_void_p()
        check(lib.ezk_prover_create(device, C.byref(self._handle)))

    def close(self) -> None:
        if getattr(self, "_handle", None) and self._handle.value:
            lib.ezk_prover_destroy(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- prove ------------------------------------------------------------------------------------------
    def _finish(self, rc: int, out, out_len) -> Proof:
        if rc != _lib.EZK_OK:
            raise ProverError(rc, lib.ezk_last_error().decode())
        try:
            return Proof(C.string_at(out, out_len.value))
        finally:
            lib.ezk_free(out)

    def prove(self, trace) -> Proof:
        """`Prover::prove(trace)`: host trace (28, n, 2) uint64 -> Proof."""
        a = _as_trace_array(trace)
        n = a.shape[1]
        cols = (C.c_void_p * TRACE_WIDTH)(*[a[c].ctypes.data for c in range(TRACE_WIDTH)])
        t = EzkTrace(C.cast(cols, C.POINTER(C.c_void_p)), TRACE_WIDTH, n)
        pi, opt = self.pub_inputs.to_c(), self.options.to_c()
        out, out_len = C.c_void_p(), C.c_size_t()
        rc = lib.ezk_prover_prove(self._handle, C.byref(t), C.byref(pi), C.byref(opt), C.byref(out), C.byref(out_len))
        return self._finish(rc, out, out_len)

    OP_COLUMNS = (0, 1, 2, 3, 4, 5, 6, 11)  # clk, op bits, chiplet flag, stack depth: built on the device from the op list

    def prove_with_ops(self, trace, op_codes, last_row=None) -> Proof:
        """`prove` with the eight bookkeeping columns generated on the device from the executed operation codes
        (`Program.ops()[0]`): the corresponding rows of `trace` are never read.  last_row defaults to trace[:, n-1]."""
        a = _as_trace_array(trace)
        n = a.shape[1]
        codes = np.ascontiguousarray(op_codes, dtype=np.uint8)
        last = np.ascontiguousarray(a[:, n - 1] if last_row is None else last_row, dtype=np.uint64).reshape(TRACE_WIDTH, 2)
        cols = (C.c_void_p * TRACE_WIDTH)(*[None if c in self.OP_COLUMNS else a[c].ctypes.data for c in range(TRACE_WIDTH)])
        t = EzkTrace(C.cast(cols, C.POINTER(C.c_void_p)), TRACE_WIDTH, n)
        ops = _lib.EzkOpList(codes.ctypes.data, len(codes), last.ctypes.data)
        pi, opt = self.pub_inputs.to_c(), self.options.to_c()
        out, out_len = C.c_void_p(), C.c_size_t()
        rc = lib.ezk_prover_prove_ops(self._handle, C.byref(t), C.byref(ops), C.byref(pi), C.byref(opt), C.byref(out),
                                      C.byref(out_len))
        return self._finish(rc, out, out_len)

    def prove_device(self, device_ptr: int, n: int) -> Proof:
        """Same with the trace already in device memory (28 contiguous columns of n elements)."""
        pi, opt = self.pub_inputs.to_c(), self.options.to_c()
        out, out_len = C.c_void_p(), C.c_size_t()
        rc = lib.ezk_prover_prove_device(self._handle, C.c_void_p(device_ptr), n, C.byref(pi), C.byref(opt),
                                         C.byref(out), C.byref(out_len))
        return self._finish(rc, out, out_len)

    # -- verify -----------------------------------------------------------------------------------------
    def verify(self, proof, min_conjectured_security: int = 95) -> None:
        """`winterfell::verify(proof, pub_inputs, &MinConjecturedSecurity(95))` (vm/src/lib.rs:91-98) against this
        prover's public inputs; raises VerifierError when the proof is rejected."""
        data = proof.to_bytes() if isinstance(proof, Proof) else bytes(proof)
        pi = self.pub_inputs.to_c()
        rc = lib.ezk_prover_verify(self._handle, data, len(data), C.byref(pi), min_conjectured_security)
        if rc == _lib.EZK_ERR_VERIFICATION:
            raise VerifierError(rc, lib.ezk_last_error().decode())
        check(rc)

    # -- multi-GPU single proof -------------------------------------------------------------------------
    def join_group(self) -> int:
        """Shards every following prove() over the ranks of the default torch.distributed group (1, 2, 4 or 8
        GPUs of one box).  Collective: every rank must call it, then call prove() with the same inputs."""
        import torch.distributed as dist
        rank, world = dist.get_rank(), dist.get_world_size()
        ident = C.create_string_buffer(128)
        if rank == 0:
            check(lib.ezk_comm_unique_id(ident))
        box = [ident.raw]
        dist.broadcast_object_list(box, src=0)
        check(lib.ezk_prover_join(self._handle, rank, world, box[0]))
        return world

    def leave_group(self) -> None:
        check(lib.ezk_prover_join(self._handle, 0, 1, bytes(128)))

    def join_local(self, group: "LocalGroup", rank: int) -> None:
        """Member `rank` of an in-process group (see LocalGroup); prove() must then run in one host thread per member."""
        check(lib.ezk_prover_join_local(self._handle, group._handle, rank))

    def leave_local(self) -> None:
        check(lib.ezk_prover_join_local(self._handle, None, 0))

    # -- timing -----------------------------------------------------------------------------------------
    def timer_start(self) -> None:
        check(lib.ezk_prover_timer_start(self._handle))

    def timer_stop(self) -> float:
        ms = C.c_float()
        check(lib.ezk_prover_timer_stop(self._handle, C.byref(ms)))
        return ms.value

    # -- introspection ----------------------------------------------------------------------------------
    def stage_times_ms(self) -> dict:
        ms = (C.c_float * len(_lib.STAGES))()
        check(lib.ezk_prover_stage_times(self._handle, ms))
        return dict(zip(_lib.STAGES, [float(x) for x in ms]))

    def artifact(self, name: str) -> bytes:
        which = _lib.ARTIFACTS.index(name)
        size = C.c_size_t()
        check(lib.ezk_prover_artifact(self._handle, which, None, 0, C.byref(size)))
        buf = C.create_string_buffer(size.value)
        check(lib.ezk_prover_artifact(self._handle, which, buf, size.value, C.byref(size)))
        return buf.raw

    # -- stage-level entry points (parity tests, stage sweep) -------------------------------------------
    def stage_lde(self, columns: np.ndarray) -> np.ndarray:
        a = np.ascontiguousarray(columns, dtype=np.uint64)
        w, n = a.shape[0], a.shape[1]
        out = np.empty((w, 8 * n, 2), dtype=np.uint64)
        check(lib.ezk_stage_lde(self._handle, a.ctypes.data, w, n, out.ctypes.data))
        return out

    def stage_ntt(self, columns: np.ndarray, inverse: bool) -> np.ndarray:
        a = np.ascontiguousarray(columns, dtype=np.uint64)
        w, n = a.shape[0], a.shape[1]
        out = np.empty_like(a)
        check(lib.ezk_stage_ntt(self._handle, a.ctypes.data, w, n, int(inverse), out.ctypes.data))
        return out

    def stage_merkle(self, table: np.ndarray) -> bytes:
        a = np.ascontiguousarray(table, dtype=np.uint64)
        w, rows = a.shape[0], a.shape[1]
        buf = C.create_string_buffer(2 * rows * 32)
        check(lib.ezk_stage_merkle(self._handle, a.ctypes.data, w, rows, buf))
        return buf.raw

    def stage_fri_fold(self, evals: np.ndarray, alpha: int) -> np.ndarray:
        a = np.ascontiguousarray(evals, dtype=np.uint64).reshape(-1, 2)
        s = a.shape[0]
        out = np.empty((s // 8, 2), dtype=np.uint64)
        al = elements_to_array([alpha])
        check(lib.ezk_stage_fri_fold(self._handle, a.ctypes.data, s, al.ctypes.data, out.ctypes.data))
        return out

    def stage_eval_frames(self, cur: np.ndarray, nxt: np.ndarray, periodic: np.ndarray, delta: int) -> np.ndarray:
        c = np.ascontiguousarray(cur, dtype=np.uint64).reshape(-1, 28, 2)
        nx = np.ascontiguousarray(nxt, dtype=np.uint64).reshape(-1, 28, 2)
        p = np.ascontiguousarray(periodic, dtype=np.uint64).reshape(-1, 9, 2)
        out = np.empty((c.shape[0], 20, 2), dtype=np.uint64)
        check(lib.ezk_stage_eval_frames(self._handle, c.ctypes.data, nx.ctypes.data, p.ctypes.data, c.shape[0], delta,
                                        out.ctypes.data))
        return out

    def stage_eval_frames_sum(self, cur: np.ndarray, nxt: np.ndarray, periodic: np.ndarray, delta: int,
                              tcoef: Sequence[int]) -> np.ndarray:
        """sum_j tcoef[j] * r_j per frame through the production (grouped, flagged) path of the constraint kernel."""
        c = np.ascontiguousarray(cur, dtype=np.uint64).reshape(-1, 28, 2)
        nx = np.ascontiguousarray(nxt, dtype=np.uint64).reshape(-1, 28, 2)
        p = np.ascontiguousarray(periodic, dtype=np.uint64).reshape(-1, 9, 2)
        tc = elements_to_array(list(tcoef))
        out = np.empty((c.shape[0], 2), dtype=np.uint64)
        check(lib.ezk_stage_eval_frames_sum(self._handle, c.ctypes.data, nx.ctypes.data, p.ctypes.data, c.shape[0], delta,
                                            tc.ctypes.data, out.ctypes.data))
        return out

    def bench_lde_merkle(self, width: int, n: int, iters: int = 3):
        a, b = C.c_float(), C.c_float()
        check(lib.ezk_bench_lde_merkle(self._handle, width, n, iters, C.byref(a), C.byref(b)))
        return a.value, b.value

    def bench_fri(self, n: int, iters: int = 3) -> float:
        a = C.c_float()
        check(lib.ezk_bench_fri(self._handle, n, iters, C.byref(a)))
        return a.value


class LocalGroup:
    """Several provers of ONE process as the ranks of a sharded proof (host-synchronised device copies instead of
    NCCL; include/ezkvm_prover.h).  `prove_sharded` runs one proof with `world` members on the given devices - all on
    cuda:0 by default, which is how the sharded pipeline is tested on a one-GPU box."""

    def __init__(self, world: int):
        self.world = world
        self._handle = C.c_void_p()
        check(lib.ezk_local_group_create(world, C.byref(self._handle)))

    def close(self) -> None:
        if getattr(self, "_handle", None) and self._handle.value:
            lib.ezk_local_group_destroy(self._handle)
            self._handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def prove_sharded_in_process(world: int, options: ProofOptions, program_hash, stack_outputs, server_key: ServerKey, trace,
                             devices: Optional[Sequence[int]] = None, device_ptr: Optional[int] = None, own_columns_only=False):
    """One proof by `world` provers of this process (threads), returns the list of the members' proof bytes.
    own_columns_only: member r is handed a trace in which every column it does not own is poisoned."""
    import threading
    a = _as_trace_array(trace)
    devices = list(devices) if devices is not None else [0] * world
    group = LocalGroup(world)
    provers = [ExecutionProver(options, program_hash, stack_outputs, server_key, device=devices[r]) for r in range(world)]
    out, err = [None] * world, [None] * world

    def work(r):
        try:
            provers[r].join_local(group, r)
            if device_ptr is not None:
                out[r] = provers[r].prove_device(device_ptr, a.shape[1]).to_bytes()
            else:
                t = a
                if own_columns_only:
                    t = a.copy()
                    for c in range(TRACE_WIDTH):
                        if c % world != r:
                            t[c] = 0xFFFFFFFFFFFFFFFF  # not even canonical: a rank that read it would fail
                out[r] = provers[r].prove(t).to_bytes()
        except Exception as e:  # noqa: BLE001 - reported to the caller below
            err[r] = e

    threads = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for p in provers:
        p.close()
    group.close()
    for e in err:
        if e is not None:
            raise e
    return out


def verify(proof, pub_inputs: PublicInputs, min_conjectured_security: int = 95, device: int = 0) -> None:
    """`winterfell::verify::<ProcessorAir, Blake3, DefaultRandomCoin<Blake3>>(proof, pub_inputs,
    &AcceptableOptions::MinConjecturedSecurity(95))` - vm/src/lib.rs:91-98, examples/linear_regression/src/main.rs:81-85."""
    with ExecutionProver(ProofOptions(), pub_inputs.program_hash, pub_inputs.stack_outputs, pub_inputs.server_key,
                         device=device) as p:
        p.verify(proof, min_conjectured_security)


class wire_compat:
    """Context manager over the process-wide `ezk_wire_compat` switches (byte-level details of the winterfell 0.9.0
    proof format that only a real reference proof can settle).  Keyword names are the struct's
    fields: ood_interleaved, remainder_low_to_high, trace_info_aux_rands_byte, first_nonce."""

    def __init__(self, **fields):
        self._fields = fields
        self._saved = None

    def __enter__(self):
        self._saved = _lib.EzkWireCompat()
        lib.ezk_get_wire_compat(C.byref(self._saved))
        cur = _lib.EzkWireCompat()
        lib.ezk_get_wire_compat(C.byref(cur))
        for k, v in self._fields.items():
            if k not in dict(_lib.EzkWireCompat._fields_) or k == "reserved":
                raise ValueError(f"unknown wire-compat switch {k!r}")
            setattr(cur, k, int(v))
        lib.ezk_set_wire_compat(C.byref(cur))
        return self

    def __exit__(self, *exc):
        lib.ezk_set_wire_compat(C.byref(self._saved))


def profile_enable(on: bool) -> None:
    lib.ezk_profile_enable(int(on))


def profile_reset() -> None:
    lib.ezk_profile_reset()


def profile_read() -> dict:
    """kernel name -> {launches, ms, algo_bytes} since the last reset (CUDA events around every launch)."""
    out = {}
    for k in range(lib.ezk_profile_kernel_count()):
        n, ms, b = C.c_uint64(), C.c_double(), C.c_uint64()
        lib.ezk_profile_read(k, C.byref(n), C.byref(ms), C.byref(b))
        if n.value:
            out[lib.ezk_profile_kernel_name(k).decode()] = {"launches": n.value, "ms": ms.value, "algo_bytes": b.value}
    return out


def device_count() -> int:
    return lib.ezk_device_count()


def kernel_launch_count() -> int:
    return int(lib.ezk_kernel_launch_count())
