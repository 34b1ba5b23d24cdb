"""Shared test cases: programs, inputs and host-built traces (the VM is host-side product code; the oracle
and the GPU prover both consume the SAME trace arrays)."""
from __future__ import annotations

from dataclasses import dataclass
from functools import lru_cache
from typing import List

import numpy as np

import encrypt_zkvm_b200 as ezk
from encrypt_zkvm_b200 import vm

LR_PROGRAM = """
# Compute
# b0 + (b1 * x1) + (b2 * x2) + (b3 * x3) + (b4 * x4)
read2
read
smul # (b1 * x1)
read2
read
smul # (b2 * x2)
add2
read2
read
smul # (b3 * x3)
add2
read2
read
smul # (b4 * x4)
add2
read
sadd
"""  # examples/linear_regression/lr.txt (program text is an input of configs[0], not code)


@dataclass
class Case:
    program: ezk.Program
    trace: np.ndarray
    program_hash: List[int]
    outputs: List[int]
    key: ezk.ServerKey


def pub_elements(program_hash, outputs) -> List[int]:
    return list(program_hash) + list(outputs)


@lru_cache(maxsize=None)
def lr_case() -> Case:
    """configs[0]: public [3,2,4,2,1], secrets enc(2,3,3,2), p=8, q=128, k=4 (examples/.../main.rs:21-47)."""
    params = ezk.LweParameters()
    key = vm.lwe_keygen(params, seed=7)
    xs = [vm.lwe_encrypt(key, v, seed=100 + i) for i, v in enumerate([2, 3, 3, 2])]
    prog = ezk.Program.compile(LR_PROGRAM)
    ex = ezk.execute(prog, ezk.ProgramInputs([3, 2, 4, 2, 1], xs, key), last_row_seed=11)
    return Case(prog, ex.trace(), prog.hash(), ex.outputs(), key)


@lru_cache(maxsize=None)
def small_case() -> Case:
    """vm/src/lib.rs:47-99 test_prove: a=1, b=3, x=enc(2); result decrypts to (a + x) * 3."""
    params = ezk.LweParameters()
    key = vm.lwe_keygen(params, seed=3)
    x = vm.lwe_encrypt(key, 2, seed=5)
    prog = ezk.Program.compile("read2\nread\nsadd\npush.1\npush.2\nadd\nsmul\n")
    ex = ezk.execute(prog, ezk.ProgramInputs([1, 3], [x], key), last_row_seed=13)
    return Case(prog, ex.trace(), prog.hash(), ex.outputs(), key)


@lru_cache(maxsize=None)
def synthetic(kind: int, log_n: int, delta: int = 16) -> Case:
    params = ezk.LweParameters(plaintext_modulus=8, ciphertext_modulus=8 * delta)
    key = ezk.ServerKey(params)
    prog, ex = ezk.synthetic_case(kind, log_n, key)
    return Case(prog, ex.trace(), prog.hash(), ex.outputs(), key)
