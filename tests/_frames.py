"""The evaluation frames of the reference's own AIR unit tests (air/src/tests/mod.rs:10-343), rebuilt in Python.

Each frame lists which of the 20 transition constraints the reference asserts to be zero on it.  Ciphertexts
that the reference draws from thread_rng are drawn from a seeded PRNG here (their values are irrelevant to the
constraint being zero).  The Rescue round uses Python big-int arithmetic with the constants of
crypto/src/rescue.rs:194-378 (include/ezkvm_rescue_constants.h).
"""
from __future__ import annotations

import random
import re
from dataclasses import dataclass
from pathlib import Path
from typing import List

M = 2**128 - 45 * 2**40 + 1
ROOT = Path(__file__).resolve().parent.parent


def _constants():
    text = (ROOT / "include" / "ezkvm_rescue_constants.h").read_text()

    def table(name):
        body = text.split("#define " + name)[1].split("#define")[0]
        return [int(lo, 16) | (int(hi, 16) << 64) for lo, hi in re.findall(r"\{0x([0-9a-f]+)ULL, 0x([0-9a-f]+)ULL\}", body)]

    inv_alpha = int(re.search(r"INV_ALPHA_LO 0x([0-9a-f]+)", text).group(1), 16) | (
        int(re.search(r"INV_ALPHA_HI 0x([0-9a-f]+)", text).group(1), 16) << 64)
    return table("EZK_RESCUE_MDS_INIT"), table("EZK_RESCUE_INV_MDS_INIT"), table("EZK_RESCUE_ARK_INIT"), inv_alpha


MDS, INV_MDS, ARK, INV_ALPHA = _constants()


def apply_round(state: List[int], op_code: int, op_value: int, step: int) -> List[int]:
    """crypto/src/rescue.rs:102-118"""
    ark = ARK[(step % 16) * 8:(step % 16) * 8 + 8]
    s = [pow(x, 3, M) for x in state]
    s = [sum(MDS[i * 4 + j] * s[j] for j in range(4)) % M for i in range(4)]
    s = [(s[i] + ark[i]) % M for i in range(4)]
    s[0] = (s[0] + op_code) % M
    s[1] = (s[1] + op_value) % M
    s = [pow(x, INV_ALPHA, M) for x in s]
    s = [sum(MDS[i * 4 + j] * s[j] for j in range(4)) % M for i in range(4)]
    return [(s[i] + ark[4 + i]) % M for i in range(4)]


def sponge_hash(ops) -> List[int]:
    """Rescue128 over (code, value) pairs (crypto/src/rescue.rs:42-60) -> [s0, s1]"""
    state = [0, 0, 0, 0]
    for step, (code, value) in enumerate(ops):
        if step % 16 < 14:
            state = apply_round(state, code, value, step)
        else:
            state[2] = state[3] = 0
    return state[:2]


@dataclass
class Frame:
    name: str
    cur: List[int]
    nxt: List[int]
    periodic: List[int]
    zero: List[int]  # constraint indices the reference asserts to be zero


def reference_frames(delta: int = 16) -> List[Frame]:
    rng = random.Random(0xA1)
    ct = lambda: [rng.randrange(M) for _ in range(5)]
    Z = lambda: [0] * 28
    P0 = [0] * 9
    out: List[Frame] = []

    cur, nxt = Z(), Z()
    cur[0], nxt[0] = 3, 4
    out.append(Frame("clock_increase", cur, nxt, P0, [0]))                      # mod.rs:10-21

    for a, b in ([0, 0], [1, 0], [0, 1]):                                       # mod.rs:23-36
        cur, nxt = Z(), Z()
        cur[4], cur[5] = a, b
        out.append(Frame(f"stack_shift_{a}{b}", cur, nxt, P0, [2]))

    for depth, opcode in zip([1, -1, 5, -5], [[0, 0, 0, 0, 1], [0, 0, 0, 1, 0], [0, 1, 0, 0, 1], [1, 1, 0, 1, 0]]):
        cur, nxt = Z(), Z()                                                     # mod.rs:38-62
        cur[1:6] = opcode
        cur[11], nxt[11] = 10, 10 + depth
        out.append(Frame(f"stack_depth_{depth}", cur, nxt, P0, [1]))

    cur, nxt = Z(), Z()                                                          # mod.rs:64-79
    cur[4] = 1
    cur[12], cur[13], nxt[12] = 4, 2, 6
    out.append(Frame("add", cur, nxt, P0, [3]))

    cur, nxt = Z(), Z()                                                          # mod.rs:81-114
    v = ct()
    cur[2] = cur[4] = 1
    cur[12] = 4
    cur[13:18] = v
    res = v[:4] + [(v[4] + delta * 4) % M]
    nxt[12:17] = res
    out.append(Frame("sadd", cur, nxt, P0, [4]))

    cur, nxt = Z(), Z()                                                          # mod.rs:116-158
    v0, v1 = ct(), ct()
    cur[1] = cur[2] = cur[4] = 1
    cur[12:17], cur[17:22] = v0, v1
    nxt[12:17] = [(a + b) % M for a, b in zip(v0, v1)]
    out.append(Frame("add2", cur, nxt, P0, [5]))

    cur, nxt = Z(), Z()                                                          # mod.rs:160-176
    cur[1] = cur[4] = 1
    cur[12], cur[13], nxt[12] = 4, 2, 8
    out.append(Frame("mul", cur, nxt, P0, [6]))

    cur, nxt = Z(), Z()                                                          # mod.rs:178-211
    v = ct()
    cur[3] = cur[4] = 1
    cur[12] = 4
    cur[13:18] = v
    nxt[12:17] = [x * 4 % M for x in v]
    out.append(Frame("smul", cur, nxt, P0, [7]))

    cur, nxt = Z(), Z()                                                          # mod.rs:213-226
    cur[5] = 1
    cur[12], nxt[13] = 4, 4
    out.append(Frame("push", cur, nxt, P0, [8]))

    cur, nxt = Z(), Z()                                                          # mod.rs:228-242
    cur[1] = cur[5] = 1
    cur[12], nxt[13] = 4, 4
    out.append(Frame("read", cur, nxt, P0, [9]))

    cur, nxt = Z(), Z()                                                          # mod.rs:244-258
    cur[2] = cur[5] = 1
    cur[12], nxt[17] = 4, 4
    out.append(Frame("read2", cur, nxt, P0, [10]))

    cur, nxt = Z(), Z()                                                          # mod.rs:260-271
    cur[12], nxt[12] = 4, 4
    out.append(Frame("noop", cur, nxt, P0, [11]))

    cur, nxt = Z(), Z()                                                          # mod.rs:273-304
    cur[5] = 1
    cur[6] = 1
    state = apply_round([0, 0, 0, 0], 16, 2, 0)
    nxt[7:11] = state
    nxt[12] = 2
    out.append(Frame("hash_round", cur, nxt, [1] + ARK[0:8], [12, 13, 14, 15]))

    cur, nxt = Z(), Z()                                                          # mod.rs:306-329
    cur[6] = 1
    cur[7:11] = [2, 4, 6, 8]
    nxt[7:11] = [2, 4, 0, 0]
    out.append(Frame("hash_copy", cur, nxt, P0, [16, 17, 18, 19]))
    return out
