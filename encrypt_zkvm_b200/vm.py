"""Host-side mirror of the reference's `vm` crate over the C ABI (the VM stays on the host).

Reference interface:
    Program::load(path) / Program::compile(source)     vm/src/program/mod.rs:29-96
    ProgramInputs::new(public, secret, &server_key)    vm/src/program/inputs.rs:11
    vm::prove(program, inputs) -> (hash, output, proof) vm/src/lib.rs:13-29
Errors carry the reference's Display text ("program error at 3: ...", "stack error at 7: ...").
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from pathlib import Path
from typing import List, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import EzkError, check, lib
from .prover import (ExecutionProver, LweParameters, Proof, ProofOptions, ServerKey, array_to_elements,
                     bytes_to_elements, elements_to_array)

OPCODES = {"noop": 0b00000, "push": 0b10000, "read": 0b10001, "read2": 0b10010, "add": 0b01000, "mul": 0b01001,
           "sadd": 0b01010, "smul": 0b01100, "add2": 0b01011}  # vm/src/processor/opcodes.rs:30-43


class ProgramError(EzkError):
    """vm/src/program/errors.rs"""


class ProcessorError(EzkError):
    """vm/src/processor/errors.rs"""


class Program:
    def __init__(self, handle: C.c_void_p):
        self._handle = handle

    @staticmethod
    def compile(source: str) -> "Program":
        h = C.c_void_p()
        rc = lib.ezk_program_compile(source.encode(), C.byref(h))
        if rc != _lib.EZK_OK:
            raise ProgramError(rc, lib.ezk_last_error().decode())
        return Program(h)

    @staticmethod
    def load(path) -> "Program":
        try:
            source = Path(path).read_text()
        except OSError as e:  # vm/src/program/mod.rs:30-33
            raise ProgramError(_lib.EZK_ERR_VM, f"program error at 0: {str(e).lower()}")
        return Program.compile(source)

    def __del__(self):
        if getattr(self, "_handle", None) and self._handle.value and lib is not None:
            lib.ezk_program_free(self._handle)
            self._handle = C.c_void_p()

    def __len__(self) -> int:
        return lib.ezk_program_len(self._handle)

    def code(self) -> List[Tuple[int, int]]:
        n = len(self)
        codes, values = (C.c_uint8 * n)(), (C.c_uint8 * n)()
        lib.ezk_program_ops(self._handle, codes, values)
        return list(zip(codes, values))

    def op_codes(self) -> np.ndarray:
        """The executed operation codes (compiler padding included) as a uint8 array: the input of
        `ExecutionProver.prove_with_ops` (device-side generation of the clk / decoder / flag / depth columns)."""
        n = len(self)
        codes = np.empty(n, dtype=np.uint8)
        lib.ezk_program_ops(self._handle, codes.ctypes.data_as(C.POINTER(C.c_uint8)), None)
        return codes

    def hash(self) -> List[int]:
        buf = C.create_string_buffer(32)
        lib.ezk_program_hash(self._handle, buf)
        return bytes_to_elements(buf.raw)

    def __str__(self) -> str:
        need = lib.ezk_program_display(self._handle, None, 0)
        buf = C.create_string_buffer(need)
        lib.ezk_program_display(self._handle, buf, need)
        return buf.value.decode()


@dataclass
class ProgramInputs:
    public: Sequence[int]
    secret: Sequence[Sequence[int]]  # ciphertexts: lwe_size elements each
    server_key: ServerKey = field(default_factory=ServerKey)


class Execution:
    """Result of Processor::run + Processor::trace (vm/src/processor/mod.rs:61-101)."""

    def __init__(self, handle: C.c_void_p):
        self._handle = handle

    def __del__(self):
        if getattr(self, "_handle", None) and self._handle.value and lib is not None:
            lib.ezk_execution_free(self._handle)
            self._handle = C.c_void_p()

    @property
    def length(self) -> int:
        return int(lib.ezk_execution_length(self._handle))

    def trace(self) -> np.ndarray:
        n = self.length
        out = np.empty((28, n, 2), dtype=np.uint64)
        for c in range(28):
            ptr = lib.ezk_execution_column(self._handle, c)
            C.memmove(out[c].ctypes.data, ptr, n * 16)
        return out

    def outputs(self) -> List[int]:
        buf = C.create_string_buffer(256)
        lib.ezk_execution_outputs(self._handle, buf)
        return bytes_to_elements(buf.raw)


def execute(program: Program, inputs: ProgramInputs, last_row_seed: int = 1) -> Execution:
    pub = bytes(bytearray(int(v) & 0xFF for v in inputs.public))
    lw = inputs.server_key.lwe_size()
    flat: List[int] = []
    for ct in inputs.secret:
        if len(ct) != lw:
            raise ValueError("ciphertext length must equal lwe_size")
        flat.extend(ct)
    sec = elements_to_array(flat) if flat else np.zeros((0, 2), dtype=np.uint64)
    h = C.c_void_p()
    rc = lib.ezk_vm_execute(program._handle, pub, len(pub), sec.ctypes.data if len(flat) else None, len(inputs.secret),
                            inputs.server_key.parameters.k, inputs.server_key.parameters.delta, last_row_seed, C.byref(h))
    if rc != _lib.EZK_OK:
        raise ProcessorError(rc, lib.ezk_last_error().decode())
    return Execution(h)


def synthetic_case(kind: int, log_n: int, server_key: ServerKey | None = None, seed: int | None = None):
    """Benchmark programs of BASELINE.md section 2 (1 scalar, 2 ciphertext, 3 mixed) -> (Program, Execution)."""
    sk = server_key or ServerKey()
    if seed is None:
        seed = 0xE2C0DE00 + log_n
    ph, eh = C.c_void_p(), C.c_void_p()
    check(lib.ezk_synthetic_case(kind, log_n, sk.parameters.k, sk.parameters.delta, seed, C.byref(ph), C.byref(eh)))
    return Program(ph), Execution(eh)


def prove(program: Program, inputs: ProgramInputs, options: ProofOptions | None = None, last_row_seed: int = 1,
          device: int = 0):
    """vm::prove (vm/src/lib.rs:13-29): run, build the trace on the host, prove on the GPU."""
    ex = execute(program, inputs, last_row_seed)
    output = ex.outputs()
    options = options or ProofOptions()
    with ExecutionProver(options, program.hash(), output, inputs.server_key, device=device) as prover:
        proof = prover.prove(ex.trace())
    return program.hash(), output, proof


# ---- LWE client side (fhe/src/server_key.rs:19-76), seeded instead of thread_rng ----
def lwe_keygen(params: LweParameters, seed: int) -> ServerKey:
    buf = np.empty((params.k, 2), dtype=np.uint64)
    lib.ezk_lwe_keygen(params.k, seed, buf.ctypes.data)
    return ServerKey(params, array_to_elements(buf))


def lwe_encrypt(key: ServerKey, value: int, seed: int) -> List[int]:
    k = key.parameters.k
    kb = elements_to_array(key.key)
    out = np.empty((k + 1, 2), dtype=np.uint64)
    lib.ezk_lwe_encrypt(kb.ctypes.data, k, key.parameters.delta, key.parameters.std, value, seed, out.ctypes.data)
    return array_to_elements(out)


def lwe_decrypt(key: ServerKey, ciphertext: Sequence[int]) -> int:
    k = key.parameters.k
    kb = elements_to_array(key.key)
    ct = elements_to_array(list(ciphertext)[:k + 1])
    return int(lib.ezk_lwe_decrypt(kb.ctypes.data, k, key.parameters.delta, ct.ctypes.data))
