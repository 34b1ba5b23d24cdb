"""Wire formats of the reference's example and `fhe` crate, host-side only.

    InputData   examples/linear_regression/src/utils.rs:36-70     server_key ‖ usize n ‖ n x FheUInt8 ‖ usize m ‖ m bytes
    OutputData  examples/linear_regression/src/utils.rs:97-128    hash (2 elements) ‖ Proof ‖ usize 16 ‖ 16 elements
    ServerKey   fhe/src/server_key.rs:131-155                      LweParameters ‖ usize k ‖ k elements
    LweParameters fhe/src/parameters.rs:24-50                      u32 p ‖ u32 q ‖ u32 delta ‖ usize k ‖ f64 std (LE)
    FheElement  fhe/src/integer.rs:30-51                           usize len ‖ len elements
Field elements are 16 little-endian bytes (canonical; non-canonical bytes are rejected like winter-math does),
`usize` is winter-utils' vint64 (`ByteWriter::write_usize`): n bytes holding ((v << 1) | 1) << (n - 1) little-endian,
n = 9 - min((leading_zeros(v) - 1) / 7, 8) bytes, with a zero first byte followed by 8 raw bytes when n = 9.
`Proof` bytes are opaque here (`Proof::write_into` = `to_bytes`); the proof is the last variable-length field before
the fixed 1 + 16 * 16 byte tail of OutputData, which is how it is delimited on reading.
Export / Import (fhe/src/lib.rs:43-88) are `to_file` / `from_file`: the serialized bytes, nothing else.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass
from pathlib import Path
from typing import List, Sequence, Tuple

from .prover import MODULUS, LweParameters, Proof, ServerKey


class DeserializationError(ValueError):
    """winterfell::DeserializationError"""


def write_usize(v: int) -> bytes:
    if not 0 <= v < 1 << 64:
        raise ValueError("usize out of range")
    zeros = 64 - v.bit_length()
    length = 9 - min(max(zeros - 1, 0) // 7, 8)
    if length == 9:
        return b"\x00" + v.to_bytes(8, "little")
    return ((((v << 1) | 1) << (length - 1)) & ((1 << 64) - 1)).to_bytes(8, "little")[:length]


class Reader:
    def __init__(self, data: bytes):
        self.data, self.pos = bytes(data), 0

    def take(self, n: int) -> bytes:
        if self.pos + n > len(self.data):
            raise DeserializationError("unexpected end of input")
        out = self.data[self.pos:self.pos + n]
        self.pos += n
        return out

    def u32(self) -> int:
        return int.from_bytes(self.take(4), "little")

    def usize(self) -> int:
        first = self.take(1)[0]
        length = 9 if first == 0 else ((first & -first).bit_length())  # trailing zeros + 1
        if length == 9:
            return int.from_bytes(self.take(8), "little")
        rest = self.take(length - 1)
        return int.from_bytes(bytes([first]) + rest, "little") >> length

    def element(self) -> int:
        v = int.from_bytes(self.take(16), "little")
        if v >= MODULUS:
            raise DeserializationError("invalid field element: value is greater than or equal to the field modulus")
        return v

    def done(self) -> bool:
        return self.pos == len(self.data)


def _elements(values: Sequence[int]) -> bytes:
    out = bytearray()
    for v in values:
        v = int(v)
        if not 0 <= v < MODULUS:
            raise ValueError("field element out of range")
        out += v.to_bytes(16, "little")
    return bytes(out)


# ---- fhe ----
def write_parameters(p: LweParameters) -> bytes:
    return (struct.pack("<III", p.plaintext_modulus, p.ciphertext_modulus, p.delta) + write_usize(p.k) + struct.pack("<d", p.std))


def read_parameters(r: Reader) -> LweParameters:
    p, q, delta = r.u32(), r.u32(), r.u32()
    k = r.usize()
    (std,) = struct.unpack("<d", r.take(8))
    params = LweParameters(p, q, k, std)
    if p == 0 or params.delta != delta:
        raise DeserializationError("inconsistent LWE parameters")
    return params


def write_server_key(key: ServerKey) -> bytes:
    elems = list(key.key or [])
    return write_parameters(key.parameters) + write_usize(len(elems)) + _elements(elems)


def read_server_key(r: Reader) -> ServerKey:
    params = read_parameters(r)
    n = r.usize()
    return ServerKey(params, [r.element() for _ in range(n)])


def write_fhe_element(ciphertext: Sequence[int]) -> bytes:
    return write_usize(len(ciphertext)) + _elements(ciphertext)


def read_fhe_element(r: Reader) -> List[int]:
    n = r.usize()
    return [r.element() for _ in range(n)]


# ---- example ----
@dataclass
class InputData:
    public_inputs: bytes
    secret_inputs: List[List[int]]
    server_key: ServerKey

    def to_bytes(self) -> bytes:
        out = bytearray(write_server_key(self.server_key))
        out += write_usize(len(self.secret_inputs))
        for ct in self.secret_inputs:
            out += write_fhe_element(ct)
        out += write_usize(len(self.public_inputs)) + bytes(self.public_inputs)
        return bytes(out)

    @staticmethod
    def from_bytes(data: bytes) -> "InputData":
        r = Reader(data)
        key = read_server_key(r)
        secrets = [read_fhe_element(r) for _ in range(r.usize())]
        public = r.take(r.usize())
        return InputData(public, secrets, key)


@dataclass
class OutputData:
    hash: Tuple[int, int]
    proof: Proof
    output: List[int]

    def to_bytes(self) -> bytes:
        if len(self.output) != 16:
            raise ValueError("expected 16 output elements")
        return _elements(self.hash) + self.proof.to_bytes() + write_usize(16) + _elements(self.output)

    @staticmethod
    def from_bytes(data: bytes) -> "OutputData":
        tail = 1 + 16 * 16
        if len(data) < 32 + tail:
            raise DeserializationError("unexpected end of input")
        r = Reader(data[:32])
        h = (r.element(), r.element())
        t = Reader(data[len(data) - tail:])
        if t.usize() != 16:
            raise DeserializationError("expected an array containing f128::BaseElement of length 16")
        out = [t.element() for _ in range(16)]
        return OutputData(h, Proof(data[32:len(data) - tail]), out)


def to_file(path, payload: bytes) -> None:
    """fhe::Export::export_to_file"""
    Path(path).write_bytes(payload)


def from_file(path) -> bytes:
    """fhe::Import::import_from_file (the caller picks the reader)"""
    return Path(path).read_bytes()
