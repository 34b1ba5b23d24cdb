"""Wire formats of the reference's example / fhe crate (encrypt_zkvm_b200/wire.py), mirroring fhe/src/tests.rs:6-60
(serialization round trips) with byte-level expectations derived from the Serializable impls
(fhe/src/parameters.rs:24-50, server_key.rs:131-155, integer.rs:30-51, examples/.../utils.rs:36-128)."""
import struct

import pytest

from encrypt_zkvm_b200 import wire
from encrypt_zkvm_b200.prover import MODULUS, LweParameters, Proof, ServerKey


@pytest.mark.parametrize("v,enc", [
    (0, "01"), (1, "03"), (5, "0b"), (16, "21"), (127, "ff"),                    # 1 byte: (v << 1) | 1
    (128, "0202"), (300, "b204"), (16383, "feff"),                                 # 2 bytes: ((v << 1) | 1) << 1
    (16384, "040002"), (2**21 - 1, "fcffff"), (2**56 - 1, "80" + "ff" * 7),
    (2**56, "00" + "00" * 7 + "01"), (2**64 - 1, "00" + "ff" * 8),                 # 9 bytes: zero marker + raw u64
])
def test_vint64_known_encodings(v, enc):
    assert wire.write_usize(v).hex() == enc
    r = wire.Reader(bytes.fromhex(enc))
    assert r.usize() == v and r.done()


def test_vint64_round_trip_all_lengths():
    for bits in range(0, 64):
        for v in {1 << bits, (1 << bits) - 1, (1 << bits) + 1}:
            if v < 1 << 64:
                b = wire.write_usize(v)
                r = wire.Reader(b)
                assert r.usize() == v and r.done(), v


def test_parameters_and_server_key_bytes():
    p = LweParameters(8, 128, 4, 2.412390240121573e-5)
    b = wire.write_parameters(p)
    assert b == struct.pack("<III", 8, 128, 16) + b"\x09" + struct.pack("<d", 2.412390240121573e-5)
    key = ServerKey(p, [1, 0, 0, 1])
    kb = wire.write_server_key(key)
    assert kb == b + b"\x09" + b"".join(int(v).to_bytes(16, "little") for v in (1, 0, 0, 1))
    back = wire.read_server_key(wire.Reader(kb))
    assert back.parameters == p and back.key == [1, 0, 0, 1]


def test_fhe_element_and_input_output_round_trip(tmp_path):
    ct = [3, MODULUS - 1, 2**100, 7, 12345678901234567890]
    b = wire.write_fhe_element(ct)
    assert b[0] == 0x0B and len(b) == 1 + 5 * 16
    assert wire.read_fhe_element(wire.Reader(b)) == ct
    key = ServerKey(LweParameters(), [0, 1, 1, 0])
    inp = wire.InputData(bytes([3, 2, 4, 2, 1]), [ct, list(reversed(ct))], key)
    blob = inp.to_bytes()
    back = wire.InputData.from_bytes(blob)
    assert back.public_inputs == inp.public_inputs and back.secret_inputs == inp.secret_inputs
    assert back.server_key.key == key.key and back.server_key.parameters == key.parameters
    out = wire.OutputData((11, 22), Proof(bytes(range(200)) * 3), list(range(16)))
    ob = out.to_bytes()
    assert ob[:16] == (11).to_bytes(16, "little") and ob[-257] == 0x21  # usize 16
    back = wire.OutputData.from_bytes(ob)
    assert back.hash == (11, 22) and back.output == list(range(16)) and back.proof.to_bytes() == out.proof.to_bytes()
    wire.to_file(tmp_path / "out.bin", ob)
    assert wire.from_file(tmp_path / "out.bin") == ob


def test_malformed_inputs_are_rejected():
    with pytest.raises(wire.DeserializationError):
        wire.read_fhe_element(wire.Reader(b"\x05" + b"\x00" * 31))          # 2 elements announced, 31 bytes
    with pytest.raises(wire.DeserializationError):
        wire.read_fhe_element(wire.Reader(b"\x03" + b"\xff" * 16))          # element >= modulus
    with pytest.raises(wire.DeserializationError):
        wire.read_parameters(wire.Reader(struct.pack("<III", 8, 128, 15) + b"\x09" + b"\x00" * 8))  # delta != q / p
    with pytest.raises(wire.DeserializationError):
        wire.OutputData.from_bytes(b"\x00" * 100)
