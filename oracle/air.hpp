// ORACLE (test infrastructure, not product): the reference AIR `ProcessorAir`, restated on the CPU.
// Every function cites the reference lines it follows.
#pragma once
#include "f128.hpp"
#include "ntt.hpp"
#include "../include/ezkvm_rescue_constants.h"
#include <vector>

namespace orc {

static const unsigned TRACE_WIDTH = 28;      // vm/src/processor/mod.rs:76-84
static const unsigned NUM_TRANSITION = 20;   // air/src/lib.rs:69-90
static const unsigned NUM_ASSERTIONS = 22;   // air/src/lib.rs:94,170-195
static const unsigned NUM_EXEMPTIONS = 2;    // air/src/lib.rs:94
static const unsigned CYCLE_LENGTH = 16;     // crypto/src/rescue.rs:14
static const unsigned NUM_PERIODIC = 9;      // air/src/lib.rs:201-205
static const unsigned NUM_COMP_COLUMNS = 7;  // SURVEY App. A.2 (AirContext::num_constraint_composition_columns)

struct Pair64 {
    uint64_t lo, hi;
};
static const Pair64 MDS_RAW[16] = {EZK_RESCUE_MDS_INIT};
static const Pair64 INV_MDS_RAW[16] = {EZK_RESCUE_INV_MDS_INIT};
static const Pair64 ARK_RAW[128] = {EZK_RESCUE_ARK_INIT};
static inline u128 c128(const Pair64& p) { return mk128(p.hi, p.lo); }

struct AirParams {
    uint32_t lwe_k;  // fhe/src/parameters.rs:8 ; lwe_size = k + 1 (server_key.rs:85-87)
    uint32_t delta;  // fhe/src/parameters.rs:7,17
};

// crypto/src/rescue.rs:146-150 (ALPHA = 3)
static inline void apply_sbox(u128 s[4]) {
    for (int i = 0; i < 4; i++) s[i] = fmul(fmul(s[i], s[i]), s[i]);
}
// crypto/src/rescue.rs:162-176 / 178-192
static inline void apply_matrix(const Pair64* mat, u128 s[4]) {
    u128 r[4] = {0, 0, 0, 0};
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) r[i] = fadd(r[i], fmul(c128(mat[i * 4 + j]), s[j]));
    for (int i = 0; i < 4; i++) s[i] = r[i];
}
static inline void apply_mds(u128 s[4]) { apply_matrix(MDS_RAW, s); }
static inline void apply_inv_mds(u128 s[4]) { apply_matrix(INV_MDS_RAW, s); }

// air/src/flags.rs:15-35: b0 = cur[5] (MSB) ... b4 = cur[1] (LSB)
struct Flags {
    u128 b0, b1, b2, b3, b4;
    u128 n0, n1, n2, n3, n4;
    explicit Flags(const u128* cur) {
        b0 = cur[5], b1 = cur[4], b2 = cur[3], b3 = cur[2], b4 = cur[1];
        n0 = fsub(1, b0), n1 = fsub(1, b1), n2 = fsub(1, b2), n3 = fsub(1, b3), n4 = fsub(1, b4);
    }
    static u128 p5(u128 a, u128 b, u128 c, u128 d, u128 e) { return fmul(fmul(fmul(fmul(a, b), c), d), e); }
    u128 shr() const { return b0; }                          // flags.rs:37-39
    u128 shl() const { return b1; }                          // flags.rs:41-43
    u128 add() const { return p5(n0, b1, n2, n3, n4); }      // flags.rs:45-47
    u128 sadd() const { return p5(n0, b1, n2, b3, n4); }     // flags.rs:49-51
    u128 add2() const { return p5(n0, b1, n2, b3, b4); }     // flags.rs:53-55
    u128 mul() const { return p5(n0, b1, n2, n3, b4); }      // flags.rs:57-59
    u128 smul() const { return p5(n0, b1, b2, n3, n4); }     // flags.rs:61-63
    u128 push() const { return p5(b0, n1, n2, n3, n4); }     // flags.rs:65-67
    u128 read() const { return p5(b0, n1, n2, n3, b4); }     // flags.rs:69-71
    u128 read2() const { return p5(b0, n1, n2, b3, n4); }    // flags.rs:73-75
    u128 noop() const { return p5(n0, n1, n2, n3, n4); }     // flags.rs:77-79
    u128 opcode() const {                                    // flags.rs:81-87
        u128 r = fmul(b0, 16);
        r = fadd(r, fmul(b1, 8));
        r = fadd(r, fmul(b2, 4));
        r = fadd(r, fmul(b3, 2));
        return fadd(r, b4);
    }
};

// air/src/lib.rs:104-168 (evaluate_transition) + air/src/constrains.rs:95-216
static inline void evaluate_transition(const u128* cur, const u128* nxt, const u128* periodic,
                                       const AirParams& ap, u128* r) {
    Flags f(cur);
    const u128* s = cur + 12;   // constrains.rs:42-48
    const u128* sn = nxt + 12;  // constrains.rs:50-56
    const unsigned lw = ap.lwe_k + 1;
    // constrains.rs:95-97
    r[0] = fsub(nxt[0], fadd(cur[0], 1));
    // constrains.rs:103-106
    {
        u128 t = fsub(fsub(nxt[11], cur[11]), f.shr());
        t = fadd(t, f.shl());
        t = fsub(t, fmul(f.read2(), 4));
        r[1] = fadd(t, fmul(f.add2(), 4));
    }
    // constrains.rs:99-101
    r[2] = fmul(f.shr(), f.shl());
    // constrains.rs:108-110
    r[3] = fmul(f.add(), fsub(sn[0], fadd(s[0], s[1])));
    // constrains.rs:112-126 with server_key.rs:78-83,104-114
    {
        u128 acc = 0;
        for (unsigned j = 0; j < lw; j++) {
            u128 trivial = (j == ap.lwe_k) ? fmul((u128)ap.delta, s[0]) : 0;
            u128 out = fadd(s[1 + j], trivial);
            acc = fadd(acc, fsub(sn[j], out));
        }
        r[4] = fmul(f.sadd(), acc);
    }
    // constrains.rs:128-144 with server_key.rs:89-102 (zip truncates to lwe_size)
    {
        u128 acc = 0;
        for (unsigned j = 0; j < lw; j++) acc = fadd(acc, fsub(sn[j], fadd(s[j], s[lw + j])));
        r[5] = fmul(f.add2(), acc);
    }
    // constrains.rs:146-148
    r[6] = fmul(f.mul(), fsub(sn[0], fmul(s[0], s[1])));
    // constrains.rs:150-164 with server_key.rs:116-124
    {
        u128 acc = 0;
        for (unsigned j = 0; j < lw; j++) acc = fadd(acc, fsub(sn[j], fmul(s[1 + j], s[0])));
        r[7] = fmul(f.smul(), acc);
    }
    // constrains.rs:166-176
    r[8] = fmul(f.push(), fsub(sn[1], s[0]));
    r[9] = fmul(f.read(), fsub(sn[1], s[0]));
    r[10] = fmul(f.read2(), fsub(sn[5], s[0]));
    // constrains.rs:178-180
    r[11] = fmul(f.noop(), fsub(sn[0], s[0]));
    // constrains.rs:182-209
    u128 hash_flag = periodic[0];
    const u128* ark = periodic + 1;
    u128 h0 = cur[6];
    {
        u128 step0[4] = {cur[7], cur[8], cur[9], cur[10]};
        apply_sbox(step0);
        apply_mds(step0);
        for (int i = 0; i < 4; i++) step0[i] = fadd(step0[i], ark[i]);
        step0[0] = fadd(step0[0], f.opcode());
        step0[1] = fadd(step0[1], fmul(sn[0], f.push()));
        u128 step1[4] = {nxt[7], nxt[8], nxt[9], nxt[10]};
        for (int i = 0; i < 4; i++) step1[i] = fsub(step1[i], ark[4 + i]);
        apply_inv_mds(step1);
        apply_sbox(step1);
        for (int i = 0; i < 4; i++) r[12 + i] = fmul(fmul(fsub(step1[i], step0[i]), hash_flag), h0);
    }
    // constrains.rs:211-216
    {
        u128 nf = fsub(1, hash_flag);
        r[16] = fmul(fmul(fsub(nxt[7], cur[7]), nf), h0);
        r[17] = fmul(fmul(fsub(nxt[8], cur[8]), nf), h0);
        r[18] = fmul(fmul(nxt[9], nf), h0);
        r[19] = fmul(fmul(nxt[10], nf), h0);
    }
}

// air/src/lib.rs:201-225 + crypto/src/rescue.rs:120-134: 9 columns of 16 values
static inline std::vector<std::vector<u128>> periodic_columns() {
    std::vector<std::vector<u128>> cols(NUM_PERIODIC, std::vector<u128>(CYCLE_LENGTH));
    for (unsigned i = 0; i < CYCLE_LENGTH; i++) {
        cols[0][i] = i < 14 ? 1 : 0;
        for (unsigned j = 0; j < 8; j++) cols[1 + j][i] = c128(ARK_RAW[i * 8 + j]);
    }
    return cols;
}
// winter-air `get_periodic_column_polys`: each column interpolated over <w_16>
static inline std::vector<std::vector<u128>> periodic_polys() {
    auto cols = periodic_columns();
    for (auto& c : cols) interpolate_poly(c);
    return cols;
}

struct Assertion {
    unsigned column;
    size_t step;
    u128 value;
};
// air/src/lib.rs:170-195, then sorted by (step, column) as winter-air `prepare_assertions` does
// (all are single assertions: stride 0).  `pub_inputs` = program_hash[2] ++ stack_outputs[16]
// (air/src/lib.rs:38-47).
static inline std::vector<Assertion> sorted_assertions(size_t n, const u128* pub_inputs) {
    size_t last = n - NUM_EXEMPTIONS;  // air/src/lib.rs:56-59
    std::vector<Assertion> a;
    a.push_back({0, 0, 0});
    a.push_back({11, 0, 0});
    for (unsigned i = 0; i < 2; i++) {
        a.push_back({i + 7, 0, 0});
        a.push_back({i + 7, last, pub_inputs[i]});
    }
    for (unsigned i = 0; i < 8; i++) {
        a.push_back({i + 12, 0, 0});
        a.push_back({i + 12, last, pub_inputs[2 + i]});
    }
    for (size_t i = 1; i < a.size(); i++)  // stable insertion sort on (step, column)
        for (size_t j = i; j > 0; j--) {
            const Assertion &x = a[j - 1], &y = a[j];
            if (x.step > y.step || (x.step == y.step && x.column > y.column)) {
                Assertion t = a[j - 1];
                a[j - 1] = a[j];
                a[j] = t;
            } else
                break;
        }
    return a;
}

}  // namespace orc
