// Device restatement of the reference AIR `ProcessorAir::evaluate_transition`
// (air/src/lib.rs:104-168, air/src/constrains.rs:95-216, air/src/flags.rs:15-91), with the Rescue round
// helpers of crypto/src/rescue.rs:146-192 and the LWE ops of fhe/src/server_key.rs:78-124 inlined.
// Generic over how frame cells are fetched (`Frame::cur(c)`, `Frame::nxt(c)`) and where the 20 constraint
// values go (`Sink::put(j, value)`), so the fused row-streaming kernel (K5) and the frame-level parity
// kernel share one body.
#pragma once
#include "../field/f128.cuh"
#include "../../../include/ezkvm_rescue_constants.h"

namespace ezk {
namespace dev {

struct AirConsts {
    uint64_t inv_mds[16][4][2];  // crypto/src/rescue.rs:216-233 (full-size constants), precomputed form (fe_pre)
};

__device__ __forceinline__ fe_pre ld_pre(const uint64_t (*c)[2]) {
    fe_pre r;
#pragma unroll
    for (int i = 0; i < 4; i++) r.w[i] = fe_make(c[i][0], c[i][1]);
    return r;
}


// MDS * v with the reference's matrix (crypto/src/rescue.rs:197-214). Entries are small in absolute value
// (positive p or M - q), so each product is a 128x32-bit multiply: r_i = sum_j sign_ij * |m_ij| * v_j.
template <class A>
__device__ __forceinline__ void rescue_mds(A& ar, const fe v[4], fe out[4]) {
    // row-major magnitudes and signs of MDS (negative entries are M - magnitude)
    const uint32_t mag[16] = {729, 1080, 390, 40, 29160, 42471, 14520, 1210, 882090, 1277640, 429429, 33880,
                              24698520, 35708310, 11935560, 925771};
    const bool neg[16] = {true, false, true, false, true, false, true, false, true, false, true, false,
                          true, false, true, false};
#pragma unroll
    for (int i = 0; i < 4; i++) {
        fe acc = fe_zero();
#pragma unroll
        for (int j = 0; j < 4; j++) {
            fe t = ar.mul_small(v[j], mag[i * 4 + j]);
            acc = neg[i * 4 + j] ? ar.sub(acc, t) : ar.add(acc, t);
        }
        out[i] = acc;
    }
}

template <class A>
__device__ __forceinline__ void rescue_inv_mds(A& ar, const AirConsts* __restrict__ k, const fe v[4], fe out[4]) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
        fe acc = fe_zero();
#pragma unroll
        for (int j = 0; j < 4; j++) {
            acc = ar.add(acc, ar.mul_pre(v[j], ld_pre(k->inv_mds[i * 4 + j])));
        }
        out[i] = acc;
        ar.checkpoint();
    }
}

// periodic: 9 values [hash_flag, ark0..ark7] for this row (air/src/lib.rs:201-205)
template <class A, class Frame, class Sink>
__device__ __forceinline__ void eval_transition(A& ar, const Frame& f, const fe* periodic, const AirConsts* __restrict__ k,
                                                uint32_t delta, Sink& sink) {
    const fe one = fe_one();
    // op bits: b0 = cur[5] (MSB) ... b4 = cur[1] (LSB)   (flags.rs:15-35)
    const fe b0 = f.cur(5), b1 = f.cur(4), b2 = f.cur(3), b3 = f.cur(2), b4 = f.cur(1);
    const fe n2 = ar.sub(one, b2);
    // Products of complemented bits expand into one product and a few differences, e.g. (1 - b3)(1 - b4) =
    // 1 - b3 - b4 + b3 b4: the same field elements as the reference's left-to-right products (flags.rs:45-79),
    // because the field operations are exact, with 6 products fewer per row.
    const fe b3b4 = ar.mul(b3, b4);
    const fe b3n4 = ar.sub(b3, b3b4), n3b4 = ar.sub(b4, b3b4);
    const fe n3n4 = ar.sub(ar.sub(one, b3), n3b4);            // 1 - b3 - (b4 - b3 b4)
    const fe b0b1 = ar.mul(b0, b1);                            // also constraint r2
    const fe n0b1 = ar.sub(b1, b0b1), b0n1 = ar.sub(b0, b0b1);
    const fe n0n1 = ar.sub(ar.sub(one, b0), n0b1);            // 1 - b0 - (b1 - b0 b1)
    ar.checkpoint();
    const fe arith = ar.mul(n0b1, n2);                    // !b0 * b1 * !b2
    const fe io = ar.mul(b0n1, n2);                       // b0 * !b1 * !b2
    // A sink that only wants sum_j coef_j * r_j (Sink::kGrouped) gets the constraints that share a selector factor
    // as coef-weighted sums times that factor: the same field value with fewer products (9 per row), since
    // multiplication distributes exactly.  add / sadd / mul are then never formed on their own.
    fe f_add = fe_zero(), f_sadd = fe_zero(), f_mul = fe_zero();
    if (!Sink::kGrouped) f_add = ar.mul(arith, n3n4), f_sadd = ar.mul(arith, b3n4), f_mul = ar.mul(arith, n3b4);
    const fe f_add2 = ar.mul(arith, b3b4);
    const fe f_smul = ar.mul(ar.mul(n0b1, b2), n3n4);
    ar.checkpoint();
    const fe f_push = ar.mul(io, n3n4), f_read = ar.mul(io, n3b4), f_read2 = ar.mul(io, b3n4);
    const fe f_noop = ar.mul(ar.mul(n0n1, n2), n3n4);

    ar.checkpoint();
    // r0: clk' - (clk + 1)                                   constrains.rs:95-97
    sink.put(ar, 0, ar.sub(f.nxt(0), ar.add(f.cur(0), one)));
    // r1: (d' - d - shr + shl) - 4*read2 + 4*add2            constrains.rs:103-106
    {
        fe t = ar.sub(ar.sub(f.nxt(11), f.cur(11)), b0);
        t = ar.add(t, b1);
        sink.put(ar, 1, ar.add(t, ar.mul_small(ar.sub(f_add2, f_read2), 4)));  // - 4 read2 + 4 add2
    }
    // r2: shr * shl                                          constrains.rs:99-101
    sink.put(ar, 2, b0b1);

    const fe s0 = f.cur(12), s1 = f.cur(13);
    const fe sn0 = f.nxt(12), sn1 = f.nxt(13);
    // r3: add * (s0' - (s0 + s1))                            constrains.rs:108-110
    // r6: mul * (s0' - s0*s1)                                constrains.rs:146-148
    const fe e3 = ar.sub(sn0, ar.add(s0, s1)), e6 = ar.sub(sn0, ar.mul(s0, s1));
    fe arith_sum = fe_zero();  // grouped: sum over the arith-selected constraints of coef * (b3, b4 selector) * expression
    if (Sink::kGrouped) {
        arith_sum = ar.add(ar.mul(n3n4, sink.scaled(ar, 3, e3)), ar.mul(n3b4, sink.scaled(ar, 6, e6)));
    } else {
        sink.put(ar, 3, ar.mul(f_add, e3));
        sink.put(ar, 6, ar.mul(f_mul, e6));
    }
    ar.checkpoint();
    // ciphertext ops: lwe_size = 5 (SURVEY 8b "constraints discovered")
    {
        // The three sums run over the same next-row cells: with N = sum_j s'_j and C1 = sum_j s_{1+j},
        //   sadd: sum_j (s'_j - ct_j) with ct = s_{1..5} + trivial(delta * s0)   = N - C1 - delta s0
        //         (constrains.rs:112-126, server_key.rs:78-83,104-114)
        //   add2: sum_j (s'_j - (s_j + s_{5+j}))                                 = N - sum_{j<10} s_j
        //         (constrains.rs:128-144, server_key.rs:89-102)
        //   smul: sum_j (s'_j - s0 s_{1+j})                                      = N - s0 C1   (ONE product)
        //         (constrains.rs:150-164, server_key.rs:116-124)
        fe nsum = fe_zero(), c1 = fe_zero(), lo10 = fe_zero();
#pragma unroll
        for (int j = 0; j < 5; j++) {
            nsum = ar.add(nsum, f.nxt(12 + j));
            c1 = ar.add(c1, f.cur(13 + j));
        }
#pragma unroll
        for (int j = 0; j < 10; j++) lo10 = ar.add(lo10, f.cur(12 + j));
        const fe acc_sadd = ar.sub(ar.sub(nsum, c1), ar.mul_small(s0, delta));
        const fe acc_add2 = ar.sub(nsum, lo10);
        const fe acc_smul = ar.sub(nsum, ar.mul(s0, c1));
        ar.checkpoint();
        if (Sink::kGrouped) {
            arith_sum = ar.add(arith_sum, ar.mul(b3n4, sink.scaled(ar, 4, acc_sadd)));
            arith_sum = ar.add(arith_sum, ar.mul(b3b4, sink.scaled(ar, 5, acc_add2)));
            sink.add(ar, ar.mul(arith, arith_sum));
        } else {
            sink.put(ar, 4, ar.mul(f_sadd, acc_sadd));
            sink.put(ar, 5, ar.mul(f_add2, acc_add2));
        }
        sink.put(ar, 7, ar.mul(f_smul, acc_smul));
    }
    ar.checkpoint();
    // r8..r11                                                 constrains.rs:166-180
    {
        const fe d1 = ar.sub(sn1, s0);
        if (Sink::kGrouped) {
            sink.add(ar, ar.mul(ar.add(sink.scaled(ar, 8, f_push), sink.scaled(ar, 9, f_read)), d1));
        } else {
            sink.put(ar, 8, ar.mul(f_push, d1));
            sink.put(ar, 9, ar.mul(f_read, d1));
        }
        sink.put(ar, 10, ar.mul(f_read2, ar.sub(f.nxt(17), s0)));
        sink.put(ar, 11, ar.mul(f_noop, ar.sub(sn0, s0)));
    }
    // Rescue round / copy                                     constrains.rs:182-216
    {
        const fe hash_flag = periodic[0];
        const fe h0 = f.cur(6);
        const fe gate_round = ar.mul(hash_flag, h0);
        const fe gate_copy = ar.mul(ar.sub(one, hash_flag), h0);
        fe h[4] = {f.cur(7), f.cur(8), f.cur(9), f.cur(10)};
        fe hn[4] = {f.nxt(7), f.nxt(8), f.nxt(9), f.nxt(10)};
        ar.checkpoint();
        // hash copy (r16..r19)
        if (Sink::kGrouped) {
            fe t = ar.add(sink.scaled(ar, 16, ar.sub(hn[0], h[0])), sink.scaled(ar, 17, ar.sub(hn[1], h[1])));
            t = ar.add(t, ar.add(sink.scaled(ar, 18, hn[2]), sink.scaled(ar, 19, hn[3])));
            sink.add(ar, ar.mul(t, gate_copy));
        } else {
            sink.put(ar, 16, ar.mul(ar.sub(hn[0], h[0]), gate_copy));
            sink.put(ar, 17, ar.mul(ar.sub(hn[1], h[1]), gate_copy));
            sink.put(ar, 18, ar.mul(hn[2], gate_copy));
            sink.put(ar, 19, ar.mul(hn[3], gate_copy));
        }
        ar.checkpoint();
        // forward half: MDS * h^3 + ark[0..4], + opcode / pushed value
        fe c[4], step0[4];
#pragma unroll
        for (int i = 0; i < 4; i++) c[i] = ar.cube(h[i]);
        ar.checkpoint();
        rescue_mds(ar, c, step0);
        ar.checkpoint();
#pragma unroll
        for (int i = 0; i < 4; i++) step0[i] = ar.add(step0[i], periodic[1 + i]);
        // opcode = 16 b0 + 8 b1 + 4 b2 + 2 b3 + b4 (flags.rs:81-87), by doubling
        fe opcode = ar.add(ar.add(b0, b0), b1);
        opcode = ar.add(ar.add(opcode, opcode), b2);
        opcode = ar.add(ar.add(opcode, opcode), b3);
        opcode = ar.add(ar.add(opcode, opcode), b4);
        step0[0] = ar.add(step0[0], opcode);
        step0[1] = ar.add(step0[1], ar.mul(sn0, f_push));
        ar.checkpoint();
        // backward half: (INV_MDS * (h' - ark[4..8]))^3
        fe d[4], step1[4];
#pragma unroll
        for (int i = 0; i < 4; i++) d[i] = ar.sub(hn[i], periodic[5 + i]);
        rescue_inv_mds(ar, k, d, step1);
        ar.checkpoint();
        if (Sink::kGrouped) {
            fe t = fe_zero();
#pragma unroll
            for (int i = 0; i < 4; i++) t = ar.add(t, sink.scaled(ar, 12 + i, ar.sub(ar.cube(step1[i]), step0[i])));
            sink.add(ar, ar.mul(t, gate_round));
        } else {
#pragma unroll
            for (int i = 0; i < 4; i++) sink.put(ar, 12 + i, ar.mul(ar.sub(ar.cube(step1[i]), step0[i]), gate_round));
        }
    }
}

}  // namespace dev
}  // namespace ezk
