"""ctypes binding of oracle/liborc.so - the CPU restatement used ONLY as a checker by tests, smoke() and
bench.py's CPU-baseline / reference legs.  Never imported by the product package."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
LIB_PATH = ROOT / "oracle" / "liborc.so"
MODULUS = 2**128 - 45 * 2**40 + 1

ART = {name: i for i, name in enumerate([
    "proof", "trace_root", "comp_root", "tcoef", "bcoef", "combined", "z", "ood_cur", "ood_next", "ood_comp", "deep_tc",
    "deep_cc", "deep_evals", "fri_roots", "fri_alphas", "remainder", "positions", "trace_lde", "comp_lde", "trace_polys",
    "comp_polys", "pow_nonce", "fri_layer_evals"])}


class OrcOptions(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in ("num_queries", "blowup", "grinding", "field_ext", "fri_fold", "fri_rem_max_deg",
                                          "lwe_k", "delta", "compat_ood_interleaved", "compat_remainder_low_to_high",
                                          "compat_trace_info_aux_rands_byte")]
    _fields_.append(("compat_first_nonce", C.c_uint64))


def default_options(delta=16, lwe_k=4, **kw) -> OrcOptions:
    o = OrcOptions(32, 8, 0, 1, 8, 127, lwe_k, delta, 1, 1, 1, 1)
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def to_arr(values) -> np.ndarray:
    out = np.empty((len(values), 2), dtype=np.uint64)
    for i, v in enumerate(values):
        out[i, 0] = int(v) & 0xFFFFFFFFFFFFFFFF
        out[i, 1] = int(v) >> 64
    return out


def from_arr(a) -> list:
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 2)
    return [int(lo) | (int(hi) << 64) for lo, hi in a]


class Oracle:
    def __init__(self, lib):
        self.lib = lib
        P = C.c_void_p
        lib.orc_last_error.restype = C.c_char_p
        lib.orc_prove.restype = P
        lib.orc_prove.argtypes = [P, C.c_size_t, P, C.POINTER(OrcOptions), C.POINTER(C.c_int), C.POINTER(C.c_double)]
        lib.orc_free_artifacts.argtypes = [P]
        lib.orc_art_size.restype = C.c_size_t
        lib.orc_art_size.argtypes = [P, C.c_int, C.c_int]
        lib.orc_art_copy.argtypes = [P, C.c_int, C.c_int, P]
        lib.orc_verify.restype = C.c_int
        lib.orc_verify.argtypes = [P, C.c_size_t, P, C.POINTER(OrcOptions), C.c_uint]
        lib.orc_validate_trace.restype = C.c_long
        lib.orc_validate_trace.argtypes = [P, C.c_size_t, P, C.c_uint32, C.c_uint32]
        lib.orc_blake3.argtypes = [P, C.c_size_t, P]
        lib.orc_merkle_rows.argtypes = [P, C.c_size_t, C.c_size_t, P, P]
        lib.orc_prove_batch.restype = C.c_size_t
        lib.orc_prove_batch.argtypes = [P, C.c_size_t, P, C.c_size_t, P, C.c_size_t]
        lib.orc_batch_root.restype = C.c_int
        lib.orc_batch_root.argtypes = [P, P, C.c_size_t, C.c_uint, P, C.c_size_t, P]
        lib.orc_evaluate_transition.argtypes = [P, P, P, C.c_uint32, C.c_uint32, P]
        lib.orc_lde_column.argtypes = [P, C.c_size_t, C.c_size_t, P, P]
        lib.orc_num_threads.restype = C.c_int
        for f in ("orc_fadd", "orc_fsub", "orc_fmul"):
            getattr(lib, f).argtypes = [P, P, P, C.c_size_t]
        lib.orc_finv.argtypes = [P, P, C.c_size_t]
        lib.orc_fexp.argtypes = [P, P, P, C.c_size_t]
        for f in ("orc_interpolate", "orc_interpolate_with_offset", "orc_forward_ntt"):
            getattr(lib, f).argtypes = [P, C.c_size_t]
        lib.orc_evaluate_with_offset.argtypes = [P, C.c_size_t, C.c_size_t, P]
        lib.orc_periodic_columns.argtypes = [P]
        lib.orc_rescue_constants.argtypes = [P, P, P]
        lib.orc_eval_horner.argtypes = [P, C.c_size_t, P, P]
        lib.orc_root_of_unity.argtypes = [C.c_uint, P]
        lib.orc_synthetic_trace.restype = C.c_int
        lib.orc_synthetic_trace.argtypes = [C.c_int, C.c_uint, C.c_uint, C.c_uint, C.c_ulonglong, P, P]

    # ---- field ----
    def _bin(self, fn, a, b):
        A, B = to_arr(a), to_arr(b)
        out = np.empty_like(A)
        fn(A.ctypes.data, B.ctypes.data, out.ctypes.data, len(a))
        return from_arr(out)

    def fadd(self, a, b): return self._bin(self.lib.orc_fadd, a, b)
    def fsub(self, a, b): return self._bin(self.lib.orc_fsub, a, b)
    def fmul(self, a, b): return self._bin(self.lib.orc_fmul, a, b)
    def fexp(self, a, e): return self._bin(self.lib.orc_fexp, a, e)

    def finv(self, a):
        A = to_arr(a)
        out = np.empty_like(A)
        self.lib.orc_finv(A.ctypes.data, out.ctypes.data, len(a))
        return from_arr(out)

    def root_of_unity(self, log_n):
        out = np.empty((1, 2), dtype=np.uint64)
        self.lib.orc_root_of_unity(log_n, out.ctypes.data)
        return from_arr(out)[0]

    # ---- hash ----
    def blake3(self, data: bytes) -> bytes:
        out = C.create_string_buffer(32)
        self.lib.orc_blake3(data, len(data), out)
        return out.raw

    # ---- transforms (numpy (n,2) uint64 in/out) ----
    def interpolate(self, a, offset=False):
        v = np.array(a, dtype=np.uint64, copy=True).reshape(-1, 2)
        (self.lib.orc_interpolate_with_offset if offset else self.lib.orc_interpolate)(v.ctypes.data, v.shape[0])
        return v

    def forward_ntt(self, a):
        v = np.array(a, dtype=np.uint64, copy=True).reshape(-1, 2)
        self.lib.orc_forward_ntt(v.ctypes.data, v.shape[0])
        return v

    def lde_column(self, col, blowup=8):
        v = np.ascontiguousarray(col, dtype=np.uint64).reshape(-1, 2)
        n = v.shape[0]
        coeffs = np.empty((n, 2), dtype=np.uint64)
        lde = np.empty((n * blowup, 2), dtype=np.uint64)
        self.lib.orc_lde_column(v.ctypes.data, n, blowup, coeffs.ctypes.data, lde.ctypes.data)
        return coeffs, lde

    # ---- Merkle ----
    def merkle_rows(self, rows: np.ndarray):
        """rows: (num_rows, width, 2) uint64 row-major -> (root, nodes bytes)"""
        r = np.ascontiguousarray(rows, dtype=np.uint64)
        num, width = r.shape[0], r.shape[1]
        root = C.create_string_buffer(32)
        nodes = C.create_string_buffer(2 * num * 32)
        self.lib.orc_merkle_rows(r.ctypes.data, num, width, root, nodes)
        return root.raw, nodes.raw

    def prove_batch(self, nodes: bytes, num_leaves: int, idx) -> bytes:
        ix = np.array(idx, dtype=np.uint64)
        cap = 1 << 20
        out = C.create_string_buffer(cap)
        n = self.lib.orc_prove_batch(nodes, num_leaves, ix.ctypes.data, len(idx), out, cap)
        return out.raw[:n]

    def batch_root(self, leaves: bytes, idx, depth: int, ser: bytes):
        ix = np.array(idx, dtype=np.uint64)
        out = C.create_string_buffer(32)
        rc = self.lib.orc_batch_root(leaves, ix.ctypes.data, len(idx), depth, ser, len(ser), out)
        return rc, out.raw

    # ---- AIR ----
    def evaluate_transition(self, cur, nxt, periodic, delta=16, lwe_k=4):
        c, n, p = to_arr(cur), to_arr(nxt), to_arr(periodic)
        out = np.empty((20, 2), dtype=np.uint64)
        self.lib.orc_evaluate_transition(c.ctypes.data, n.ctypes.data, p.ctypes.data, lwe_k, delta, out.ctypes.data)
        return from_arr(out)

    def periodic_columns(self):
        out = np.empty((9 * 16, 2), dtype=np.uint64)
        self.lib.orc_periodic_columns(out.ctypes.data)
        v = from_arr(out)
        return [v[p * 16:(p + 1) * 16] for p in range(9)]

    def rescue_constants(self):
        mds, inv, ark = (np.empty((k, 2), dtype=np.uint64) for k in (16, 16, 128))
        self.lib.orc_rescue_constants(mds.ctypes.data, inv.ctypes.data, ark.ctypes.data)
        return from_arr(mds), from_arr(inv), from_arr(ark)

    def validate_trace(self, trace: np.ndarray, pub18, delta=16, lwe_k=4) -> int:
        t = np.ascontiguousarray(trace, dtype=np.uint64)
        cols = (C.c_void_p * 28)(*[t[c].ctypes.data for c in range(28)])
        pub = to_arr(pub18)
        return int(self.lib.orc_validate_trace(cols, t.shape[1], pub.ctypes.data, lwe_k, delta))

    # ---- input generator of the CPU legs (oracle/tracegen.cpp) ----
    def synthetic_trace(self, kind: int, log_n: int, seed: int | None = None, delta=16, lwe_k=4):
        """BASELINE.md's synthetic case built without the product library: (trace (28, n, 2) uint64, pub18 ints)."""
        seed = 0xE2C0DE00 + log_n if seed is None else seed
        trace = np.empty((28, 1 << log_n, 2), dtype=np.uint64)
        pub = np.empty((18, 2), dtype=np.uint64)
        if self.lib.orc_synthetic_trace(kind, log_n, lwe_k, delta, seed, trace.ctypes.data, pub.ctypes.data) != 0:
            raise RuntimeError("oracle trace generator failed")
        return trace, from_arr(pub)

    # ---- prover / verifier ----
    def prove(self, trace: np.ndarray, pub18, options: OrcOptions | None = None):
        t = np.ascontiguousarray(trace, dtype=np.uint64)
        cols = (C.c_void_p * 28)(*[t[c].ctypes.data for c in range(28)])
        pub = to_arr(pub18)
        opt = options or default_options()
        err, secs = C.c_int(), C.c_double()
        h = self.lib.orc_prove(cols, t.shape[1], pub.ctypes.data, C.byref(opt), C.byref(err), C.byref(secs))
        if not h:
            raise RuntimeError(f"oracle prove failed ({err.value}): {self.lib.orc_last_error().decode()}")
        return Artifacts(self, h, secs.value)

    def verify(self, proof: bytes, pub18, options: OrcOptions | None = None, min_security=95) -> int:
        pub = to_arr(pub18)
        opt = options or default_options()
        return int(self.lib.orc_verify(proof, len(proof), pub.ctypes.data, C.byref(opt), min_security))


class Artifacts:
    def __init__(self, oracle: Oracle, handle, seconds: float):
        self._o, self._h, self.seconds = oracle, handle, seconds

    def __del__(self):
        if self._h:
            self._o.lib.orc_free_artifacts(self._h)
            self._h = None

    def raw(self, name: str, sub: int = 0) -> bytes:
        n = self._o.lib.orc_art_size(self._h, ART[name], sub)
        buf = C.create_string_buffer(n)
        self._o.lib.orc_art_copy(self._h, ART[name], sub, buf)
        return buf.raw

    def elements(self, name: str, sub: int = 0):
        b = self.raw(name, sub)
        return [int.from_bytes(b[i:i + 16], "little") for i in range(0, len(b), 16)]

    def array(self, name: str, sub: int = 0) -> np.ndarray:
        return np.frombuffer(self.raw(name, sub), dtype=np.uint64).reshape(-1, 2)

    @property
    def proof(self) -> bytes:
        return self.raw("proof")

    @property
    def positions(self):
        return list(np.frombuffer(self.raw("positions"), dtype=np.uint64))


_cached = None


def load() -> Oracle:
    global _cached
    if _cached is None:
        if not LIB_PATH.exists():
            from encrypt_zkvm_b200.build import build_oracle
            build_oracle()
        _cached = Oracle(C.CDLL(str(LIB_PATH)))
    return _cached
