"""ctypes binding of libezkvm.so (the C ABI in include/ezkvm_prover.h).

The library is the product: there is no Python/CPU implementation behind it.  Importing this module fails
loudly when the shared object is missing; compute calls fail with EZK_ERR_NO_DEVICE when no GPU is visible.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path


_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "libezkvm.so"

EZK_OK = 0
EZK_ERR_INVALID_ARGUMENT = -1
EZK_ERR_UNSUPPORTED_FIELD_EXTENSION = -2
EZK_ERR_CONSTRAINT_DEGREE = -3
EZK_ERR_DEEP_DEGREE = -4
EZK_ERR_NO_DEVICE = -5
EZK_ERR_CUDA = -6
EZK_ERR_VM = -7
EZK_ERR_INTERNAL = -8
EZK_ERR_VERIFICATION = -9

STAGES = ["upload", "trace_lde", "trace_commit", "constraints", "composition", "deep", "fri", "queries"]
ARTIFACTS = ["trace_root", "constraint_root", "combined", "ood_trace", "ood_constraints", "deep_evals", "fri_roots",
             "remainder", "positions", "trace_lde", "constraint_lde", "trace_polys"]


class EzkOptions(C.Structure):
    _fields_ = [("num_queries", C.c_uint32), ("blowup_factor", C.c_uint32), ("grinding_factor", C.c_uint32),
                ("field_extension", C.c_uint32), ("fri_folding_factor", C.c_uint32),
                ("fri_remainder_max_degree", C.c_uint32)]


class EzkPublicInputs(C.Structure):
    _fields_ = [("program_hash", (C.c_uint8 * 16) * 2), ("stack_outputs", (C.c_uint8 * 16) * 16),
                ("lwe_k", C.c_uint32), ("lwe_delta", C.c_uint32)]


class EzkWireCompat(C.Structure):
    _fields_ = [("ood_interleaved", C.c_uint32), ("remainder_low_to_high", C.c_uint32),
                ("trace_info_aux_rands_byte", C.c_uint32), ("reserved", C.c_uint32), ("first_nonce", C.c_uint64)]


class EzkOpList(C.Structure):
    _fields_ = [("codes", C.c_void_p), ("count", C.c_uint64), ("last_row", C.c_void_p)]


class EzkTrace(C.Structure):
    _fields_ = [("columns", C.POINTER(C.c_void_p)), ("width", C.c_uint32), ("length", C.c_uint64)]


# every symbol declared in include/ezkvm_prover.h: name -> (restype, argtypes)
_P = C.c_void_p
SIGNATURES = {
    "ezk_last_error": (C.c_char_p, []),
    "ezk_version": (C.c_char_p, []),
    "ezk_device_count": (C.c_int, []),
    "ezk_kernel_launch_count": (C.c_uint64, []),
    "ezk_free": (None, [_P]),
    "ezk_default_options": (None, [C.POINTER(EzkOptions)]),
    "ezk_get_wire_compat": (None, [C.POINTER(EzkWireCompat)]),
    "ezk_set_wire_compat": (None, [C.POINTER(EzkWireCompat)]),
    "ezk_selftest_copy_pool": (C.c_int, [C.c_uint32, C.c_size_t]),
    "ezk_selftest_shard_layout": (C.c_int, [C.c_uint32, C.c_uint32, C.c_uint64]),
    "ezk_selftest_host_field": (C.c_int, [_P, _P, C.c_size_t, _P]),
    "ezk_selftest_launch_groups": (C.c_int, [C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, _P, _P, _P]),
    "ezk_prover_create": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "ezk_prover_destroy": (None, [_P]),
    "ezk_prover_prove": (C.c_int, [_P, C.POINTER(EzkTrace), C.POINTER(EzkPublicInputs), C.POINTER(EzkOptions),
                                   C.POINTER(_P), C.POINTER(C.c_size_t)]),
    "ezk_prover_prove_ops": (C.c_int, [_P, C.POINTER(EzkTrace), C.POINTER(EzkOpList), C.POINTER(EzkPublicInputs),
                                       C.POINTER(EzkOptions), C.POINTER(_P), C.POINTER(C.c_size_t)]),
    "ezk_prover_prove_device": (C.c_int, [_P, _P, C.c_uint64, C.POINTER(EzkPublicInputs), C.POINTER(EzkOptions),
                                          C.POINTER(_P), C.POINTER(C.c_size_t)]),
    "ezk_prove": (C.c_int, [C.POINTER(EzkTrace), C.POINTER(EzkPublicInputs), C.POINTER(EzkOptions), C.POINTER(_P),
                            C.POINTER(C.c_size_t)]),
    "ezk_prover_verify": (C.c_int, [_P, _P, C.c_size_t, C.POINTER(EzkPublicInputs), C.c_uint32]),
    "ezk_comm_unique_id": (C.c_int, [_P]),
    "ezk_prover_join": (C.c_int, [_P, C.c_int, C.c_int, _P]),
    "ezk_local_group_create": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "ezk_local_group_destroy": (None, [_P]),
    "ezk_prover_join_local": (C.c_int, [_P, _P, C.c_int]),
    "ezk_prover_stage_times": (C.c_int, [_P, C.POINTER(C.c_float)]),
    "ezk_prover_timer_start": (C.c_int, [_P]),
    "ezk_prover_timer_stop": (C.c_int, [_P, C.POINTER(C.c_float)]),
    "ezk_profile_enable": (None, [C.c_int]),
    "ezk_profile_reset": (None, []),
    "ezk_profile_kernel_count": (C.c_int, []),
    "ezk_profile_kernel_name": (C.c_char_p, [C.c_int]),
    "ezk_profile_read": (None, [C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_double), C.POINTER(C.c_uint64)]),
    "ezk_prover_artifact": (C.c_int, [_P, C.c_int, _P, C.c_size_t, C.POINTER(C.c_size_t)]),
    "ezk_stage_lde": (C.c_int, [_P, _P, C.c_uint32, C.c_uint64, _P]),
    "ezk_stage_merkle": (C.c_int, [_P, _P, C.c_uint32, C.c_uint64, _P]),
    "ezk_stage_fri_fold": (C.c_int, [_P, _P, C.c_uint64, _P, _P]),
    "ezk_stage_eval_frames": (C.c_int, [_P, _P, _P, _P, C.c_uint32, C.c_uint32, _P]),
    "ezk_stage_eval_frames_sum": (C.c_int, [_P, _P, _P, _P, C.c_uint32, C.c_uint32, _P, _P]),
    "ezk_stage_ntt": (C.c_int, [_P, _P, C.c_uint32, C.c_uint64, C.c_int, _P]),
    "ezk_bench_lde_merkle": (C.c_int, [_P, C.c_uint32, C.c_uint64, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "ezk_bench_fri": (C.c_int, [_P, C.c_uint64, C.c_int, C.POINTER(C.c_float)]),
    "ezk_program_compile": (C.c_int, [C.c_char_p, C.POINTER(_P)]),
    "ezk_program_free": (None, [_P]),
    "ezk_program_len": (C.c_size_t, [_P]),
    "ezk_program_ops": (None, [_P, _P, _P]),
    "ezk_program_hash": (None, [_P, _P]),
    "ezk_program_display": (C.c_size_t, [_P, C.c_char_p, C.c_size_t]),
    "ezk_vm_execute": (C.c_int, [_P, _P, C.c_size_t, _P, C.c_size_t, C.c_uint32, C.c_uint32, C.c_uint64, C.POINTER(_P)]),
    "ezk_execution_free": (None, [_P]),
    "ezk_execution_length": (C.c_uint64, [_P]),
    "ezk_execution_column": (_P, [_P, C.c_uint32]),
    "ezk_execution_outputs": (None, [_P, _P]),
    "ezk_synthetic_case": (C.c_int, [C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint64, C.POINTER(_P), C.POINTER(_P)]),
    "ezk_lwe_keygen": (None, [C.c_uint32, C.c_uint64, _P]),
    "ezk_lwe_encrypt": (None, [_P, C.c_uint32, C.c_uint32, C.c_double, C.c_uint8, C.c_uint64, _P]),
    "ezk_lwe_decrypt": (C.c_uint8, [_P, C.c_uint32, C.c_uint32, _P]),
}


class EzkError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"[ezk {code}] {message}")
        self.code = code
        self.message = message


def _load() -> C.CDLL:
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m encrypt_zkvm_b200.build` "
            "(nvcc, sm_100a). This package has no CPU fallback.")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header and library disagree
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


def check(rc: int) -> None:
    if rc != EZK_OK:
        raise EzkError(rc, lib.ezk_last_error().decode())
