// Batched column NTT / coset-LDE kernels (K1, K2).  See ntt.cuh.
//
// Decomposition (mixed-radix, natural order in and out, no bit-reversal pass over HBM):
//   n = n_1 * n_2 (* n_3);  input index m = m_1 + n_1 m_2 + n_1 n_2 m_3.
//   pass i (i = p..2, "strided"): size-n_i transform over m_i at stride n_1..n_{i-1}, in place, followed by
//       the twiddle w_{N_i}^{lo * j_i}, N_i = n_1..n_i, lo = index below digit i.  A CTA handles n_i x lanes
//       elements (lanes = 8 or 4 adjacent lo values = 128/64-byte global segments).
//   pass 1 ("final"): size-n_1 transform over contiguous runs, written out of place to the natural output
//       index j = j_p + n_p (j_{p-1} + ... n_2 j_1); a CTA handles `lanes` runs whose outputs are adjacent
//       (or LDE cosets of one run), so global writes are again 64..128-byte segments.
// Inside a tile: radix-8 decimation-in-frequency steps on REGISTERS (each thread holds 8 elements = three
// butterfly stages between shared-memory exchanges; a leading radix-2/4 step absorbs log2(n_i) mod 3); the
// first step reads global memory straight into registers and the last one stores straight from registers, so a
// 2^9-point tile crosses shared memory twice instead of nine times.  The integer pipe, not HBM, bounds these
// kernels: ~0.5 modmul (58 instr) + 1 add/sub (~10 instr) per element per stage.
#include "ntt.cuh"
#include "../field/f128.cuh"
#include "../field/f128_host.h"
#include "../common.h"
#include <map>
#include <vector>

namespace ezk {

using namespace dev;

namespace {

constexpr uint32_t kTileElems = 4096;  // field elements staged per CTA (64 KiB): S points x lanes
constexpr int kDefaultOrderLog = 1;         // default of EZK_NTT_ORDER (-1 = column-major); measured: profiles/r02_ntt_cta_order_ab.log
constexpr int kDefaultFinalOrderLog = -1;   // default of EZK_NTT_FINAL_ORDER (the final pass)

// lanes = independent transforms handled side by side by one CTA: enough of them to fill the 4096-element tile
// (so that all 256 threads have a group of 8 in every step) and to make global segments 64..512 bytes
__host__ __device__ __forceinline__ uint32_t lanes_log_for(uint32_t log_s) {
    const uint32_t want = log_s >= 12 ? 0u : 12u - log_s;
    return want < 2 ? 2u : (want > 5 ? 5u : want);
}
// threads of a CTA: one per group of 8 elements, at most 256
inline unsigned threads_for(uint32_t log_s, uint32_t lanes_log, unsigned max_threads) {
    const unsigned groups = (1u << (log_s - 3)) << lanes_log;
    const unsigned t = groups < max_threads ? groups : max_threads;
    return t < 32 ? 32 : t;
}

// w_8^1..3 for the in-register 8-point butterflies: [0] forward, [1] inverse, in precomputed form
// (c_w8pre[d][k][i] = w_8^(+-k) * 2^(32 i) mod M, see fe_mul_pre).  The direction is a template parameter, so every
// limb is a constant-bank operand of the IMAD.WIDE that uses it: no load, no register.
__constant__ uint4 c_w8pre[2][4][4];

template <int INV, int K>
__device__ __forceinline__ fe_pre w8_pre() {
    fe_pre r;
#pragma unroll
    for (int i = 0; i < 4; i++) r.w[i] = fe_from(c_w8pre[INV][K][i]);
    return r;
}

// compact per-size twiddle tables: tw[(1 << k) + e] = w_{2^k}^e, e < 2^k, k <= kTwMaxLog
constexpr uint32_t kTwMaxLog = 11;

// In-tile twiddles in precomputed form: a second set of compact tables holds every twiddle as 4 x 16 bytes
// (w * 2^(32 i) mod M) and the in-tile twiddle products use fe_mul_pre like the 8-point DFT constants do - 42 instead
// of 58 instructions per product for 4 x the table bytes (256 KiB per direction, L1/L2 resident) and up to 16 instead
// of 4 registers per twiddle in flight.  EZK_NTT_PRE_TWIDDLES: 0 off (default), 1 both passes, 2 the final pass only.
// Measured A/B on one B200, same box, 2^20 rows (profiles/r02_ntt_pre_twiddles_ab.log): 0 -> final pass 7.04 ms,
// proof 26.7 ms, e2e 29.2 ms; 2 -> 7.40 / 27.1 / 29.7 (the twiddle loads become the top stall: long_scoreboard 3.3
// warps per issue in ncu); 1 -> strided pass 10.05 -> 11.44 ms.  (The round-1 end-of-round probe that showed 2 ahead
// did not reproduce.)  Which passes use them is the Pass type's kPreTw.
#ifndef EZK_NTT_PRE_TWIDDLES
#define EZK_NTT_PRE_TWIDDLES 0
#endif
// EZK_NTT_PRE_PASS_TABLE=1 (measured in round 1, slower: 512 MiB of table traffic for the 2^20 LDE): the full
// inter-pass twiddle tables of the strided passes hold their entries in precomputed form as well (64 bytes per entry)
// and the output factor of a strided pass becomes a fe_mul_pre.
#ifndef EZK_NTT_PRE_PASS_TABLE
#define EZK_NTT_PRE_PASS_TABLE 0
#endif
constexpr uint32_t kPassTableWords = EZK_NTT_PRE_PASS_TABLE ? 4 : 1;  // 16-byte words per table entry

template <bool PRE>
struct Tw;
template <>
struct Tw<false> {
    typedef fe type;
    static __device__ __forceinline__ fe at(const uint4* __restrict__ tw, uint32_t log_size, uint32_t e) {
        return fe_ldg(tw + (1u << log_size) + e);
    }
    template <class A>
    static __device__ __forceinline__ fe mul(A& ar, const fe& x, const fe& w) {
        return ar.mul(x, w);
    }
};
template <>
struct Tw<true> {
    typedef fe_pre type;
    static __device__ __forceinline__ fe_pre at(const uint4* __restrict__ tw, uint32_t log_size, uint32_t e) {
        return fe_pre_ldg(tw + 4 * ((1u << log_size) + e));
    }
    template <class A>
    static __device__ __forceinline__ fe mul(A& ar, const fe& x, const fe_pre& w) {
        return ar.mul_pre(x, w);
    }
};

// b^e from a two-level table with the policy's product (see fe_tab_pow)
template <class A>
__device__ __forceinline__ fe tab_pow(A& ar, const uint4* __restrict__ tab, uint32_t e) {
    const uint32_t lo = e & (EZK_TAB_SIZE - 1), hi = e >> EZK_TAB_BITS;
    const fe a = fe_ldg(tab + lo), b = fe_ldg(tab + EZK_TAB_SIZE + hi);
    if (hi == 0) return a;
    if (lo == 0) return b;
    return ar.mul(a, b);
}
template <class A>
__device__ __forceinline__ fe root_pow(A& ar, const uint4* __restrict__ tab, uint32_t log_n, uint64_t e) {
    return tab_pow(ar, tab, (uint32_t)(e & ((1ull << log_n) - 1)) << (EZK_ROOT_LOG - log_n));
}

template <class A>
__device__ __forceinline__ void bf(A& ar, fe& a, fe& b) {
    fe s = ar.add(a, b), d = ar.sub(a, b);
    a = s, b = d;
}
template <bool PRE, class A>
__device__ __forceinline__ void bf_w(A& ar, fe& a, fe& b, const typename Tw<PRE>::type& w) {
    fe s = ar.add(a, b), d = ar.sub(a, b);
    a = s, b = Tw<PRE>::mul(ar, d, w);
}
template <class A>
__device__ __forceinline__ void bf_pre(A& ar, fe& a, fe& b, const fe_pre& w) {
    fe s = ar.add(a, b), d = ar.sub(a, b);
    a = s, b = ar.mul_pre(d, w);
}
__device__ __forceinline__ void swap_fe(fe& a, fe& b) {
    fe t = a;
    a = b, b = t;
}

// x[k] <- sum_p x[p] w_8^(pk): three decimation-in-frequency stages on registers, 5 constant multiplications
template <int INV, class A>
__device__ __forceinline__ void dft8(A& ar, fe (&x)[8]) {
    bf(ar, x[0], x[4]);
    bf_pre(ar, x[1], x[5], w8_pre<INV, 1>());
    bf_pre(ar, x[2], x[6], w8_pre<INV, 2>());
    bf_pre(ar, x[3], x[7], w8_pre<INV, 3>());
    bf(ar, x[0], x[2]);
    bf_pre(ar, x[1], x[3], w8_pre<INV, 2>());
    bf(ar, x[4], x[6]);
    bf_pre(ar, x[5], x[7], w8_pre<INV, 2>());
    bf(ar, x[0], x[1]);
    bf(ar, x[2], x[3]);
    bf(ar, x[4], x[5]);
    bf(ar, x[6], x[7]);
    // register r now holds frequency bitrev3(r): put frequency k into x[k]
    swap_fe(x[1], x[4]);
    swap_fe(x[3], x[6]);
}

// 4-point DFT of (m0..m3): X_k = sum_j m_j w_4^(jk), returned in natural order in the same registers
template <int INV, class A>
__device__ __forceinline__ void dft4(A& ar, fe& m0, fe& m1, fe& m2, fe& m3) {
    bf(ar, m0, m2);
    bf_pre(ar, m1, m3, w8_pre<INV, 2>());  // w_4
    bf(ar, m0, m1);
    bf(ar, m2, m3);
    swap_fe(m1, m2);
}

// One radix-2^a step (a = 1, 2, 3) of a decimation-in-frequency transform of size 2^log_cur on the 8 registers a
// thread holds: slot p (p = 0..7) is position q + p * 2^(log_cur-3) of the sub-transform.  On return x[p'] is the
// value to put back into slot p' (in place), already multiplied by its twiddle w_{2^log_cur}^(q' m').
template <int INV, bool PRE, class A>
__device__ __forceinline__ void radix_step(A& ar, fe (&x)[8], uint32_t a, uint32_t log_cur, uint32_t q,
                                           const uint4* __restrict__ tw) {
    const uint32_t eighth = 1u << (log_cur - 3);
    if (a == 3) {
        dft8<INV>(ar, x);
        if (log_cur > 3 && q != 0) {
#pragma unroll
            for (int k = 1; k < 8; k++) x[k] = Tw<PRE>::mul(ar, x[k], Tw<PRE>::at(tw, log_cur, q * k));
        }
    } else if (a == 2) {
        // two 4-point butterflies: h = p & 1 selects q' = q + h * eighth, m = p >> 1 is the digit
        dft4<INV>(ar, x[0], x[2], x[4], x[6]);
        dft4<INV>(ar, x[1], x[3], x[5], x[7]);
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const uint32_t qq = q + h * eighth;
            if (qq != 0) {
#pragma unroll
                for (int m = 1; m < 4; m++) x[2 * m + h] = Tw<PRE>::mul(ar, x[2 * m + h], Tw<PRE>::at(tw, log_cur, qq * m));
            }
        }
    } else {
        // four 2-point butterflies: h = p & 3, digit m = p >> 2 (w^0 = 1 for q' = 0)
#pragma unroll
        for (int h = 0; h < 4; h++) bf_w<PRE>(ar, x[h], x[h + 4], Tw<PRE>::at(tw, log_cur, q + h * eighth));
    }
}

// digits of a position, most significant first: a1 bits, then 3-bit digits; returns the frequency index
// (digits in reverse order, each digit's bits kept in order)
__device__ __forceinline__ uint32_t digit_reverse(uint32_t i, uint32_t log_s, uint32_t a1) {
    uint32_t j = 0, shift = 0, rem = log_s, a = a1;
    while (rem) {
        rem -= a;
        j |= ((i >> rem) & ((1u << a) - 1)) << shift;
        shift += a;
        a = 3;
    }
    return j;
}

struct StepCtx {
    uint32_t log_s, lanes_log, a, a1, log_cur, first, last;
    const uint4* tw;
};

// One group of 8 elements of one step: gather (global memory through P.load on the first step, shared memory
// otherwise), butterflies, and on the last step the pass's output factor (P.finish).  Returns true when a FAST
// computation hit a rare tail and must be redone.
template <bool FAST, int INV, class Pass>
__device__ __forceinline__ bool group_compute(const Pass& P, const uint4* tile, const StepCtx& c, uint32_t lane, uint32_t i0,
                                              uint32_t q, uint32_t jg, fe (&x)[8]) {
    Arith<FAST, Pass::kLean, Pass::kLeanMul> ar;
    const uint32_t sh = c.log_cur - 3;
    // slot p of the group sits at a fixed stride from slot 0: one base address per group, p * step per slot
    if (c.first) {
        const typename Pass::In in = P.begin_in(lane, i0, sh);
#pragma unroll
        for (int p = 0; p < 8; p++) x[p] = P.load(ar, in, p);
    } else {
        const uint4* t0 = tile + (i0 << c.lanes_log) + lane;
        const uint32_t step = (1u << sh) << c.lanes_log;
#pragma unroll
        for (int p = 0; p < 8; p++) x[p] = fe_load(t0 + p * step);
    }
    radix_step<INV, Pass::kPreTw>(ar, x, c.a, c.log_cur, q, c.tw);
    if (c.last) {
        const typename Pass::Out out = P.begin_out(lane, jg, c.log_s);
#pragma unroll
        for (int p = 0; p < 8; p++) x[p] = P.finish(ar, out, p, x[p]);
    }
    return ar.tainted();
}

template <class Pass>
__device__ __forceinline__ void group_store(const Pass& P, uint4* tile, const StepCtx& c, uint32_t lane, uint32_t i0,
                                            uint32_t jg, const fe (&x)[8]) {
    if (c.last) {
        const typename Pass::Out out = P.begin_out(lane, jg, c.log_s);
#pragma unroll
        for (int p = 0; p < 8; p++) P.store(out, p, x[p]);
    } else {
        uint4* t0 = tile + (i0 << c.lanes_log) + lane;
        const uint32_t step = (1u << (c.log_cur - 3)) << c.lanes_log;
#pragma unroll
        for (int p = 0; p < 8; p++) fe_store(t0 + p * step, x[p]);
    }
}

// the rare redo: exact arithmetic, computes and stores the group (kept out of line, off the hot path's registers)
template <int INV, class Pass>
__device__ __noinline__ void group_redo_exact(const Pass& P, uint4* tile, const StepCtx& c, uint32_t lane, uint32_t i0,
                                              uint32_t q, uint32_t jg) {
    fe x[8];
    group_compute<false, INV>(P, tile, c, lane, i0, q, jg, x);
    group_store(P, tile, c, lane, i0, jg, x);
}

// Size-2^log_s transforms of `lanes` interleaved sequences by one CTA.  Every thread keeps 8 elements in
// registers per step; steps exchange through shared memory (tile[i * lanes + lane], lanes fastest so that a
// quarter-warp always touches 8 consecutive 16-byte words); the first step reads global memory through
// P.load(lane, i) and the last writes through P.store(lane, j, v) with j the natural frequency index.
template <int INV, class Pass>
__device__ __forceinline__ void tile_transform(const Pass& P, uint4* tile, uint32_t log_s, uint32_t lanes_log,
                                               const uint4* __restrict__ tw) {
    const uint32_t steps = (log_s + 2) / 3;
    StepCtx c;
    c.log_s = log_s, c.lanes_log = lanes_log, c.tw = tw;
    c.a1 = log_s - 3 * (steps - 1);
    c.log_cur = log_s;
    const uint32_t groups = (1u << (log_s - 3)) << lanes_log;
    const uint32_t lane_mask = (1u << lanes_log) - 1;
    for (uint32_t step = 0; step < steps; step++) {
        c.a = step == 0 ? c.a1 : 3;
        c.first = step == 0, c.last = step + 1 == steps;
        const uint32_t sh = c.log_cur - 3;
        for (uint32_t u = threadIdx.x; u < groups; u += blockDim.x) {
            const uint32_t lane = u & lane_mask, g = u >> lanes_log;
            const uint32_t q = g & ((1u << sh) - 1);
            const uint32_t i0 = ((g >> sh) << c.log_cur) | q;  // slot 0; slot p adds p << sh
            // last step (sh == 0): i = 8 g + p and the last digit is the most significant part of the frequency
            const uint32_t jg = c.last ? digit_reverse(i0, log_s, c.a1) : 0;
            fe x[8];
            if (group_compute<true, INV>(P, tile, c, lane, i0, q, jg, x))
                group_redo_exact<INV>(P, tile, c, lane, i0, q, jg);
            else
                group_store(P, tile, c, lane, i0, jg, x);
        }
        if (!c.last) __syncthreads();
        c.log_cur -= c.a;
    }
}

// ---------------------------------------------------------------------------------------------------------
// pass i >= 2 ("strided"): size-n_i transform over digit i at stride n_1..n_{i-1}, in place, then the twiddle
// w_{N_i}^(lo * j_i).  A tile is n_i points x `lanes` adjacent `lo` values.
struct StridedArgs {
    const uint4* src;
    uint4* dst;
    uint64_t src_pitch, dst_pitch;  // per grid.y column, in elements
    uint32_t log_stride, log_s;     // stride = lo range, S = n_i
    uint32_t lanes_log;
    uint32_t coset_first;           // LDE first pass: column y reads coefficient column y / nc, scaled by w_L^(c*m),
    uint32_t cs_log, cs_base, cs_step;  //   nc = 1 << cs_log cosets per column, c = cs_base + cs_step * (y % nc)
    uint32_t log_L;
    uint32_t inv;
    uint32_t order_log;             // CTA order, see ntt_strided_pass
    const uint4* roots;             // two-level 2^28-th root table (forward or inverse)
    const uint4* tw;                // compact per-size tables (same direction)
    const uint4* big;               // optional full inter-pass twiddle table (see build_pass_table), else nullptr
};

struct StridedPass {
    // arithmetic encodings, see Arith in f128.cuh: this pass (latency-bound on its strided loads and table reads, and
    // at the register limit) gains nothing from the shorter ones (10.04 vs 10.02 ms) and loses with the spills the
    // carry-free product flag brings (10.43 ms)
    static constexpr bool kLean = false;
    static constexpr int kLeanMul = 0;
    static constexpr bool kPreTw = EZK_NTT_PRE_TWIDDLES == 1;
    const uint4* src;
    uint4* dst;
    const uint4* roots;
    const uint4* big;  // big[(coset << log_N) + (j << log_stride) + lo] = w_L^(lo (8 j + coset)) (w_N^(lo j) for plain passes)
    uint64_t base;
    uint32_t lo0, log_stride, log_N, coset, log_L;
    uint32_t hi_shift;  // log_stride - (log_L - EZK_TAB_BITS), see load
    // LDE first pass: the coset factor w_L^(c * idx), idx = lo + stride * m, splits into w_L^(c * stride * m)
    // (applied on load: exponent with >= 14 trailing zero bits in the 2^28 table -> one load, no product) and
    // w_L^(c * lo), constant along the transform, which is folded into the output twiddle's exponent:
    // w_N^(lo j) w_L^(c lo) = w_L^(lo (8 j + coset)).
    struct In {
        const uint4* ptr;   // slot 0
        uint64_t step;      // elements between slots
        uint32_t m0, mstep; // transform index of slot 0 / between slots (coset factor)
    };
    struct Out {
        uint4* ptr;
        const uint4* tab;   // slot 0 of the inter-pass twiddle table (same slot stride as the data), or nullptr
        uint64_t step;
        uint32_t lo, j0, jstep;
    };
    __device__ __forceinline__ In begin_in(uint32_t lane, uint32_t i0, uint32_t sh) const {
        In in;
        in.ptr = src + base + lane + ((uint64_t)i0 << log_stride);
        in.step = 1ull << (sh + log_stride);
        in.m0 = i0, in.mstep = 1u << sh;
        return in;
    }
    template <class A>
    __device__ __forceinline__ fe load(A& ar, const In& in, int p) const {
        fe v = fe_load(in.ptr + p * in.step);
        const uint32_t m = in.m0 + p * in.mstep;
        // w_L^(c m stride): stride >= L / 2^14 (run_strided checks it), so the exponent is a multiple of 2^14 in the
        // two-level 2^28-th root table and the factor is one load from its upper level
        if (coset != 0 && m != 0) v = ar.mul(v, fe_ldg(roots + EZK_TAB_SIZE + (((coset * m) << hi_shift) & (EZK_TAB_SIZE - 1))));
        return v;
    }
    __device__ __forceinline__ Out begin_out(uint32_t lane, uint32_t jg, uint32_t log_s) const {
        Out out;
        const uint64_t at = ((uint64_t)jg << log_stride) + lo0 + lane;
        out.ptr = dst + base - lo0 + at;
        out.step = 1ull << (log_s - 3 + log_stride);
        out.tab = big + kPassTableWords * (((uint64_t)coset << log_N) + at);  // used when big != nullptr
        out.lo = lo0 + lane, out.j0 = jg, out.jstep = 1u << (log_s - 3);
        return out;
    }
    template <class A>
    __device__ __forceinline__ fe finish(A& ar, const Out& out, int p, fe v) const {
#if EZK_NTT_PRE_PASS_TABLE
        if (big) return ar.mul_pre(v, fe_pre_ldg(out.tab + 4 * p * out.step));
#else
        if (big) return ar.mul(v, fe_ldg(out.tab + p * out.step));
#endif
        const uint64_t ex = (uint64_t)out.lo * (((uint64_t)(out.j0 + p * out.jstep) << (log_L - log_N)) + coset);
        return ex != 0 ? ar.mul(v, root_pow(ar, roots, log_L, ex)) : v;
    }
    __device__ __forceinline__ void store(const Out& out, int p, fe v) const { fe_store(out.ptr + p * out.step, v); }
};

template <int THREADS, int MINB, int INV>
__global__ void __launch_bounds__(THREADS, MINB) ntt_strided_pass(StridedArgs a) {
    extern __shared__ uint4 tile[];
    const uint32_t lanes = 1u << a.lanes_log;
    const uint32_t tiles_per_hi = (uint32_t)((1ull << a.log_stride) >> a.lanes_log);
    // Which (tile, column) this CTA takes.  CTAs are dispatched in the order of x + gridDim.x * y.  Column-major order
    // (tile = x, column = y: the ~450 resident CTAs work on two columns) re-reads from DRAM whatever the columns share:
    // the inter-pass twiddle table (every column needs all of it) and, in an LDE, the coefficient column that the 8
    // coset copies read.  Tile-major order (order_log = k): 2^k adjacent tiles of ALL columns back to back, so the
    // resident CTAs share one slab of the table and of the coefficients and both come from DRAM once; k > 0 keeps
    // neighbouring 64-byte segments of a row (adjacent tiles) in flight together.
    // (k = log2(gridDim.x) is the column-major order again: one formula, no branch)
    const uint32_t lin = blockIdx.x + gridDim.x * blockIdx.y;
    const uint32_t per = gridDim.y << a.order_log;  // CTAs per group of 2^k tiles
    const uint32_t grp = lin / per, r = lin - grp * per;
    const uint32_t by = r >> a.order_log;
    const uint32_t bx = (grp << a.order_log) | (r & ((1u << a.order_log) - 1));
    StridedPass P;
    P.lo0 = (bx % tiles_per_hi) * lanes;
    const uint64_t hi = bx / tiles_per_hi;
    P.base = P.lo0 + (hi << (a.log_stride + a.log_s));
    uint32_t col = by;
    if (a.coset_first) {
        // coset-major column order: the CTAs of all columns of one coset run back to back, so that (in column-major
        // order) that coset's slice of the inter-pass twiddle table is read from DRAM once and then served by L2
        const uint32_t real_cols = gridDim.y >> a.cs_log;
        const uint32_t ci = by / real_cols;
        col = ((by - ci * real_cols) << a.cs_log) | ci;
        P.coset = a.cs_base + a.cs_step * (col & ((1u << a.cs_log) - 1));
        P.src = a.src + (uint64_t)(col >> a.cs_log) * a.src_pitch;
    } else {
        P.coset = 0;
        P.src = a.src + (uint64_t)col * a.src_pitch;
    }
    P.dst = a.dst + (uint64_t)col * a.dst_pitch;
    P.roots = a.roots;
    P.big = a.big;
    P.log_stride = a.log_stride, P.log_N = a.log_stride + a.log_s;
    P.log_L = a.coset_first ? a.log_L : P.log_N;  // plain passes: exponent lo * j of w_N
    P.hi_shift = a.log_stride + EZK_TAB_BITS - a.log_L;  // used with coset_first only
    tile_transform<INV>(P, tile, a.log_s, a.lanes_log, a.tw);
}

// ---------------------------------------------------------------------------------------------------------
// pass 1 ("final"): size-n_1 transform over contiguous runs, written out of place to the natural output index.
struct FinalArgs {
    const uint4* src;
    uint4* dst;
    uint64_t src_pitch, dst_pitch;  // per column, in elements
    uint32_t log_n, log_s;          // S = n_1
    uint32_t passes;                // p in {1,2,3}
    uint32_t log_top;               // log2(n_p), the most significant storage digit (p >= 2)
    uint32_t mode;                  // 0 plain, 1 LDE from tmp (lanes = cosets), 2 LDE direct from coefficients (p = 1)
    uint32_t lanes_log;             // log2(runs handled side by side by one CTA)
    uint32_t lanes_c_log;           // LDE: the low lanes_c_log lane bits select a coset, the bits above an adjacent run (j_p)
    uint32_t log_L;
    uint32_t inv;
    uint32_t cs_log, cs_base, cs_step;  // LDE: 1 << cs_log cosets are computed, coset k = cs_base + cs_step * k
    uint32_t order_log;             // CTA order as in StridedArgs (the columns of a plain transform share the scale table)
    const uint4* roots;
    const uint4* tw;
    const uint4* off_tab;
    const uint4* scale_tab;  // optional: scale_tab[m] = cvec[0] * o^m (see NttTables::ScaleTable)
    NttScale scale;
};

struct FinalPass {
    static constexpr bool kLean = true;
    static constexpr int kLeanMul = 1;
    static constexpr bool kPreTw = EZK_NTT_PRE_TWIDDLES != 0;
    const FinalArgs* a;
    const uint4* src;   // column base (mode 0, 2) or coset-0 base of this column group (mode 1)
    uint4* dst;
    uint64_t run0, run_step, out_base;
    uint32_t log_H, coset0;
    struct In {
        const uint4* ptr;
        uint32_t step, m0, coset;
    };
    struct Out {
        uint4* ptr;
        uint64_t step, index;  // elements between slots / output index of slot 0 (for the scaling)
    };
    __device__ __forceinline__ In begin_in(uint32_t lane, uint32_t i0, uint32_t sh) const {
        In in;
        in.step = 1u << sh, in.m0 = i0, in.coset = 0;
        if (a->mode == 0) {
            in.ptr = src + ((run0 + lane * run_step) << a->log_s) + i0;
        } else if (a->mode == 1) {
            // coset0 + lc = index into this prover's coset list; the coset itself is cs_base + cs_step * index
            const uint32_t lc = lane & ((1u << a->lanes_c_log) - 1), lj = lane >> a->lanes_c_log;
            in.ptr = src + (uint64_t)(coset0 + lc) * a->src_pitch + ((run0 + lj * run_step) << a->log_s) + i0;
        } else {
            in.ptr = src + i0;
            in.coset = a->cs_base + a->cs_step * (coset0 + lane);
        }
        return in;
    }
    template <class A>
    __device__ __forceinline__ fe load(A& ar, const In& in, int p) const {
        fe v = fe_load(in.ptr + p * in.step);
        if (in.coset != 0) v = ar.mul(v, root_pow(ar, a->roots, a->log_L, (uint64_t)in.coset * (in.m0 + p * in.step)));
        return v;
    }
    __device__ __forceinline__ Out begin_out(uint32_t lane, uint32_t jg, uint32_t log_s) const {
        Out out;
        if (a->mode == 0) {
            out.index = ((uint64_t)jg << log_H) + out_base + lane;
            out.step = 1ull << (log_s - 3 + log_H);
        } else {
            // packed row order: the rows of the cosets this prover computes, ascending (all 8 cosets: the natural order)
            const uint32_t lc = lane & ((1u << a->lanes_c_log) - 1), lj = lane >> a->lanes_c_log;
            out.index = ((((uint64_t)jg << log_H) + out_base + lj) << a->cs_log) + coset0 + lc;
            out.step = 1ull << (log_s - 3 + log_H + a->cs_log);
        }
        out.ptr = dst + out.index;
        return out;
    }
    template <class A>
    __device__ __forceinline__ fe finish(A& ar, const Out& o, int p, fe v) const {
        if (a->mode == 0 && a->scale.enabled) {
            const uint64_t out = o.index + p * o.step;
            if (a->scale_tab) return ar.mul(v, fe_ldg(a->scale_tab + out));
            const uint32_t ci = (uint32_t)(out >> a->scale.chunk_shift);
            fe c = fe_make(a->scale.cvec[ci][0], a->scale.cvec[ci][1]);
            if (a->scale.use_offset) c = ar.mul(c, tab_pow(ar, a->off_tab, (uint32_t)out));
            v = ar.mul(v, c);
        }
        return v;
    }
    __device__ __forceinline__ void store(const Out& o, int p, fe v) const { fe_store(o.ptr + p * o.step, v); }
};

template <int THREADS, int MINB, int INV>
__global__ void __launch_bounds__(THREADS, MINB) ntt_final_pass(const __grid_constant__ FinalArgs a) {
    extern __shared__ uint4 tile[];
    const uint32_t lanes = 1u << a.lanes_log;
    // tile-major CTA order, see ntt_strided_pass
    const uint32_t lin = blockIdx.x + gridDim.x * blockIdx.y;
    const uint32_t per = gridDim.y << a.order_log;
    const uint32_t grp = lin / per, r = lin - grp * per;
    const uint32_t col = r >> a.order_log;
    const uint32_t bx = (grp << a.order_log) | (r & ((1u << a.order_log) - 1));
    FinalPass P;
    P.a = &a;
    P.log_H = a.log_n - a.log_s;  // log2(runs per column)
    P.coset0 = 0;
    P.dst = a.dst + (uint64_t)col * a.dst_pitch;
    if (a.mode == 0) {
        P.src = a.src + (uint64_t)col * a.src_pitch;
        if (a.passes == 1) {
            P.run0 = 0, P.run_step = 0, P.out_base = 0;
        } else {
            // run index hi' = rest + (H / n_p) * j_p ; the tile owns j_p = jp0 .. jp0+lanes-1 for one `rest`
            const uint32_t log_rest = P.log_H - a.log_top;
            const uint32_t tiles_per_rest = (1u << a.log_top) >> a.lanes_log;
            const uint64_t rest = bx / tiles_per_rest;
            const uint32_t jp0 = (bx % tiles_per_rest) * lanes;
            P.run0 = rest + ((uint64_t)jp0 << log_rest);
            P.run_step = 1ull << log_rest;
            P.out_base = (rest << a.log_top) + jp0;  // j_p + n_p * rest
        }
    } else {
        // LDE: the lanes are cosets (all of this prover's, or a part of them) and, when a prover computes fewer cosets
        // than a tile has lanes (multi-GPU), adjacent runs j_p as in the plain transform: either way the lanes of a
        // tile write adjacent rows of the packed output
        const uint32_t lj_log = a.lanes_log - a.lanes_c_log;
        const uint32_t halves_log = a.cs_log - a.lanes_c_log;
        const uint64_t hp = bx >> halves_log;  // run tile
        P.coset0 = (bx & ((1u << halves_log) - 1)) << a.lanes_c_log;
        P.src = a.mode == 1 ? a.src + (((uint64_t)col * a.src_pitch) << a.cs_log) : a.src + (uint64_t)col * a.src_pitch;
        if (a.passes <= 1) {
            P.run0 = hp, P.run_step = 0, P.out_base = 0;
        } else {
            // consecutive CTAs take consecutive `rest`, i.e. read adjacent runs (measured: 6 % faster at 2^20 than the
            // order that makes their outputs adjacent)
            const uint32_t log_rest = P.log_H - a.log_top;
            const uint64_t rest = hp & ((1ull << log_rest) - 1);
            const uint32_t jp0 = (uint32_t)(hp >> log_rest) << lj_log;
            P.run0 = rest + ((uint64_t)jp0 << log_rest);
            P.run_step = 1ull << log_rest;
            P.out_base = (rest << a.log_top) + jp0;  // j_p + n_p * rest
        }
    }
    tile_transform<INV>(P, tile, a.log_s, a.lanes_log, a.tw);
}

// n = 2 or 4 (single run, fewer points than one 8-element register group): direct summation, one thread per output
__global__ void ntt_tiny_kernel(const uint4* src, uint64_t src_pitch, uint4* dst, uint64_t dst_pitch, uint32_t log_n,
                                const uint4* roots, const uint4* off_tab, NttScale scale) {
    const uint32_t n = 1u << log_n, j = threadIdx.x, col = blockIdx.x;
    if (j >= n) return;
    const uint4* x = src + (uint64_t)col * src_pitch;
    fe acc = fe_zero();
    for (uint32_t m = 0; m < n; m++) acc = fe_add(acc, fe_mul(fe_load(x + m), fe_root_pow(roots, log_n, (uint64_t)m * j)));
    if (scale.enabled) {
        const uint32_t ci = (uint32_t)((uint64_t)j >> scale.chunk_shift);
        fe c = fe_make(scale.cvec[ci][0], scale.cvec[ci][1]);
        if (scale.use_offset) c = fe_mul(c, fe_tab_pow(off_tab, j));
        acc = fe_mul(acc, c);
    }
    fe_store(dst + (uint64_t)col * dst_pitch + j, acc);
}

// Composition columns from per-coset interpolations (multi-GPU).  For each of the 8 LDE cosets c, Y_c is the
// plain inverse transform (no 1/n) of the n constraint evaluations on that coset.  With H'_j[m] = 3^m H_j[m]:
//     Y_c[m] = n w_L^(cm) sum_j (3^n w_8^c)^j H'_j[m]   =>   H'_j[m] = s_j sum_c w_8^(-cj) w_L^(-cm) Y_c[m],  s_j = 3^(-nj) / L
// A rank handles the coefficient slice m in [m0, m0 + count): y[c * count + (m - m0)] in, out[j * count + (m - m0)] for
// j < 7 out; *flag |= 1 when an H'_7[m] is non-zero (composition degree >= 7n).
struct RecombineConsts {
    uint64_t s[8][2];
};
template <class AR>
__device__ __forceinline__ bool recombine_point(const uint4* __restrict__ y, uint64_t count, uint64_t at, const fe& u,
                                                const RecombineConsts& k, fe (&x)[8]) {
    AR ar;
    fe p = u;
#pragma unroll
    for (uint32_t c = 0; c < 8; c++) {
        fe v = fe_ldg(y + c * count + at);
        if (c > 0) {
            v = ar.mul(v, p);
            if (c < 7) p = ar.mul(p, u);
        }
        x[c] = v;
    }
    dft8<1>(ar, x);
#pragma unroll
    for (uint32_t j = 0; j < 8; j++) x[j] = ar.mul(x[j], fe_make(k.s[j][0], k.s[j][1]));
    return ar.tainted();
}
__global__ void __launch_bounds__(128) composition_recombine_kernel(const uint4* __restrict__ y, uint32_t log_n, uint64_t m0,
                                                                   uint64_t count, const uint4* __restrict__ root_inv,
                                                                   RecombineConsts k, uint4* __restrict__ out,
                                                                   uint32_t* __restrict__ flag) {
    const uint64_t at = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (at >= count) return;
    const fe u = fe_root_pow(root_inv, log_n + 3, m0 + at);  // w_L^-m
    fe x[8];
    if (recombine_point<Arith<true>>(y, count, at, u, k, x)) recombine_point<Arith<false>>(y, count, at, u, k, x);
#pragma unroll
    for (uint32_t j = 0; j < 7; j++) fe_store(out + j * count + at, x[j]);
    if (!fe_is_zero(x[7])) atomicOr(flag, 1u);
}

size_t tile_bytes(uint32_t log_s, uint32_t lanes_log) { return (((size_t)1 << log_s) << lanes_log) * sizeof(uint4); }

struct Plan {
    uint32_t passes;
    uint32_t log_d[3];  // log n_1, n_2, n_3
};

Plan make_plan(uint32_t log_n, int max_tile_log) {
    Plan p{};
    const uint32_t T = (uint32_t)max_tile_log;
    if (log_n <= T) {
        p.passes = 1, p.log_d[0] = log_n;
    } else if (log_n <= 2 * T) {
        p.passes = 2, p.log_d[0] = (log_n + 1) / 2, p.log_d[1] = log_n / 2;
    } else {
        p.passes = 3;
        p.log_d[0] = (log_n + 2) / 3;
        uint32_t rest = log_n - p.log_d[0];
        p.log_d[1] = (rest + 1) / 2, p.log_d[2] = rest / 2;
    }
    return p;
}

// launch shapes: 0 = 256 threads x 3 CTAs/SM (<= 85 registers), 1 = 512 threads x 2 CTAs/SM (<= 64 registers)
int g_variant = -1, g_variant_final = 0;
template <int T, int B, int INV>
void set_smem_attr() {
    const int bytes = (int)(kTileElems * 16);
    EZK_CUDA(cudaFuncSetAttribute(ntt_strided_pass<T, B, INV>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    EZK_CUDA(cudaFuncSetAttribute(ntt_final_pass<T, B, INV>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
}
void ensure_smem_attr() {
    if (g_variant >= 0) return;
    set_smem_attr<256, 3, 0>(), set_smem_attr<256, 3, 1>();
    set_smem_attr<512, 2, 0>(), set_smem_attr<512, 2, 1>();
    set_smem_attr<256, 2, 0>(), set_smem_attr<256, 2, 1>();
    const char* env = getenv("EZK_NTT_VARIANT");
    g_variant = env ? atoi(env) : 0;
    const char* envf = getenv("EZK_NTT_FINAL_VARIANT");  // launch shape of the final pass alone (measurement knob)
    g_variant_final = envf ? atoi(envf) : g_variant;
}
template <int T, int B>
void launch_strided_tb(dim3 grid, size_t smem, cudaStream_t s, const StridedArgs& a) {
    const unsigned threads = threads_for(a.log_s, a.lanes_log, T);
    if (a.inv)
        ntt_strided_pass<T, B, 1><<<grid, threads, smem, s>>>(a);
    else
        ntt_strided_pass<T, B, 0><<<grid, threads, smem, s>>>(a);
}
template <int T, int B>
void launch_final_tb(dim3 grid, size_t smem, cudaStream_t s, const FinalArgs& a) {
    const unsigned threads = threads_for(a.log_s, a.lanes_log, T);
    if (a.inv)
        ntt_final_pass<T, B, 1><<<grid, threads, smem, s>>>(a);
    else
        ntt_final_pass<T, B, 0><<<grid, threads, smem, s>>>(a);
}
void launch_strided(dim3 grid, size_t smem, cudaStream_t s, const StridedArgs& a) {
    if (g_variant == 1)
        launch_strided_tb<512, 2>(grid, smem, s, a);
    else if (g_variant == 2)
        launch_strided_tb<256, 2>(grid, smem, s, a);
    else
        launch_strided_tb<256, 3>(grid, smem, s, a);
}
void launch_final(dim3 grid, size_t smem, cudaStream_t s, const FinalArgs& a) {
    if (g_variant_final == 1)
        launch_final_tb<512, 2>(grid, smem, s, a);
    else if (g_variant_final == 2)
        launch_final_tb<256, 2>(grid, smem, s, a);
    else
        launch_final_tb<256, 3>(grid, smem, s, a);
}

// Full inter-pass twiddle table of one strided pass: entry (c << log_N) + (j << log_stride) + lo =
// w_L^(lo (8 j + c)) for the LDE's first pass (c = coset, 8 N entries) or w_N^(lo j) for a plain pass (N entries).
// It replaces the two-level lookup (2 loads + 1 product per element) by one coalesced load; the kernels are
// integer-pipe bound, so trading a product for 16 more bytes of (L2 / HBM) traffic per element pays.
__global__ void build_pass_table_kernel(const uint4* __restrict__ roots, uint32_t log_N, uint32_t log_stride, uint32_t log_L,
                                        uint32_t ncosets, uint4* __restrict__ out) {
    const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= ((uint64_t)ncosets << log_N)) return;
    const uint64_t c = e >> log_N, idx = e & ((1ull << log_N) - 1);
    const uint64_t j = idx >> log_stride, lo = idx & ((1ull << log_stride) - 1);
#if EZK_NTT_PRE_PASS_TABLE
    fe w = fe_root_pow(roots, log_L, lo * ((j << (log_L - log_N)) + c));
    const fe two32 = fe_make(1ull << 32, 0);
    for (int i = 0; i < 4; i++) {  // w * 2^(32 i) mod M, canonical
        fe_store(out + 4 * e + i, w);
        w = fe_mul(w, two32);
    }
#else
    fe_store(out + e, fe_root_pow(roots, log_L, lo * ((j << (log_L - log_N)) + c)));
#endif
}

const uint4* pass_table(const NttTables& t, cudaStream_t s, bool inverse, bool coset, uint32_t log_N, uint32_t log_stride,
                        uint32_t log_L) {
    const uint64_t entries = (coset ? 8ull : 1ull) << log_N;
    const uint64_t table_bytes = entries * 16 * kPassTableWords;
    if (!t.big_tables || table_bytes > t.big_table_limit_bytes) return nullptr;
    const uint64_t key = ((uint64_t)log_N << 32) | ((uint64_t)log_stride << 16) | ((uint64_t)coset << 1) | (uint64_t)inverse;
    auto it = t.big_tables->find(key);
    if (it != t.big_tables->end()) return it->second;
    uint4* d = nullptr;
    if (cudaMalloc(&d, table_bytes) != cudaSuccess) {
        cudaGetLastError();
        (*t.big_tables)[key] = nullptr;  // not enough memory: keep using the two-level tables
        return nullptr;
    }
    {
        LaunchScope ls(s, K_NTT_STRIDED, table_bytes);
        build_pass_table_kernel<<<(unsigned)((entries + 255) / 256), 256, 0, s>>>(inverse ? t.root_inv : t.root_fwd, log_N, log_stride,
                                                                                 coset ? log_L : log_N, coset ? 8 : 1, d);
    }
    EZK_CUDA(cudaGetLastError());
    (*t.big_tables)[key] = d;
    return d;
}

__global__ void build_scale_table_kernel(const uint4* __restrict__ off_tab, fe c, uint64_t n, uint4* __restrict__ out) {
    const uint64_t m = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m < n) fe_store(out + m, fe_mul(c, fe_tab_pow(off_tab, (uint32_t)m)));
}

// S[m] = c * o^m for m < 2^log_n, cached for the latest (length, constant); nullptr when it cannot be used
const uint4* scale_table(const NttTables& t, cudaStream_t s, uint32_t log_n, const NttScale& sc) {
    NttTables::ScaleTable* st = t.scale_table;
    if (!st || sc.use_offset != 1 || sc.chunk_shift < 32 || log_n > 27 || (16ull << log_n) > t.big_table_limit_bytes) return nullptr;
    if (st->d && st->log_n == log_n && st->c[0] == sc.cvec[0][0] && st->c[1] == sc.cvec[0][1]) return st->d;
    if (st->d) {
        EZK_CUDA(cudaStreamSynchronize(s));  // the old table may still be read by a kernel in flight on this stream
        cudaFree(st->d);
        st->d = nullptr;
    }
    if (cudaMalloc(&st->d, 16ull << log_n) != cudaSuccess) {
        cudaGetLastError();
        st->d = nullptr;
        return nullptr;
    }
    const uint64_t n = 1ull << log_n;
    {
        LaunchScope ls(s, K_NTT_FINAL, n * 16);
        build_scale_table_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(t.off_fwd, fe_make(sc.cvec[0][0], sc.cvec[0][1]), n, st->d);
    }
    EZK_CUDA(cudaGetLastError());
    st->log_n = log_n, st->c[0] = sc.cvec[0][0], st->c[1] = sc.cvec[0][1];
    return st->d;
}

// CTA order of a strided pass (see ntt_strided_pass): tile-major in groups of 2^k adjacent tiles, k from
// EZK_NTT_ORDER (read per launch: the A/B tool switches it inside one process; -1 = column-major grid order)
uint32_t order_log_for(dim3 grid, int k) {
    uint32_t log_gx = 0;
    while ((1u << log_gx) < grid.x) log_gx++;
    if ((1u << log_gx) != grid.x) throw CudaError("ntt: tiles per column must be a power of two");
    if ((uint64_t)grid.x * grid.y >= (1ull << 32)) throw CudaError("ntt: grid too large");
    return k < 0 || (uint32_t)k > log_gx ? log_gx : (uint32_t)k;  // log2(grid.x) = column-major
}
uint32_t strided_order_log(dim3 grid) {
    int k = kDefaultOrderLog;
    if (const char* e = getenv("EZK_NTT_ORDER")) k = atoi(e);
    return order_log_for(grid, k);
}

uint32_t final_order_log(dim3 grid) {
    int k = kDefaultFinalOrderLog;
    if (const char* e = getenv("EZK_NTT_FINAL_ORDER")) k = atoi(e);
    return order_log_for(grid, k);
}

// strided passes p..2 over `ncols` arrays of n elements: the first pass reads `first_src` and writes `buf`,
// later passes run in place in `buf`.  With `coset` the first pass reads coefficient column y/8 and applies the
// coset factor w_L^(c*m), c = y%8 (LDE).
int run_strided(const NttTables& t, cudaStream_t s, const Plan& pl, uint32_t log_n, bool inverse, const uint4* first_src,
                uint64_t first_src_pitch, const CosetSet* cs, uint4* buf, uint64_t buf_pitch, uint32_t ncols, uint32_t log_L) {
    int launches = 0;
    uint32_t log_stride = log_n;
    for (int i = (int)pl.passes - 1; i >= 1; i--) {
        log_stride -= pl.log_d[i];
        StridedArgs a{};
        bool first = (i == (int)pl.passes - 1) && first_src != nullptr;
        a.src = first ? first_src : buf;
        a.src_pitch = first ? first_src_pitch : buf_pitch;
        a.dst = buf, a.dst_pitch = buf_pitch;
        a.log_stride = log_stride, a.log_s = pl.log_d[i];
        a.lanes_log = lanes_log_for(a.log_s);
        if (a.lanes_log > log_stride) a.lanes_log = log_stride;
        a.coset_first = (first && cs) ? 1 : 0;
        if (a.coset_first && log_stride + EZK_TAB_BITS < log_L) throw CudaError("ntt: coset factor is not in the upper table level");
        if (cs) a.cs_log = cs->count_log, a.cs_base = cs->base, a.cs_step = cs->step;
        a.log_L = log_L;
        a.inv = inverse ? 1 : 0;
        a.roots = inverse ? t.root_inv : t.root_fwd;
        a.tw = StridedPass::kPreTw ? (inverse ? t.twp_inv : t.twp_fwd) : (inverse ? t.tw_inv : t.tw_fwd);
        a.big = pass_table(t, s, inverse, a.coset_first != 0, log_stride + a.log_s, log_stride, log_L);
        dim3 grid((unsigned)(((uint64_t)1 << (log_n - pl.log_d[i])) >> a.lanes_log), ncols);
        a.order_log = strided_order_log(grid);
        {
            // compulsory traffic: every element of every column read once and written once (the 8 coset copies of
            // an LDE first pass share one read of the coefficient column)
            const uint64_t elems = (uint64_t)ncols << log_n;
            LaunchScope ls(s, K_NTT_STRIDED, (a.coset_first ? (elems >> cs->count_log) + elems : 2 * elems) * 16);
            launch_strided(grid, tile_bytes(a.log_s, a.lanes_log), s, a);
        }
        EZK_CUDA(cudaGetLastError());
        launches++;
    }
    return launches;
}

}  // namespace

void ntt_tables_init(NttTables& t) {
    auto upload = [](const std::vector<Fp>& h) {
        uint4* d = nullptr;
        EZK_CUDA(cudaMalloc(&d, h.size() * sizeof(uint4)));
        EZK_CUDA(cudaMemcpy(d, h.data(), h.size() * sizeof(uint4), cudaMemcpyHostToDevice));
        return d;
    };
    auto build = [&](Fp base) {
        std::vector<Fp> h(2 * EZK_TAB_SIZE);
        Fp acc(1);
        for (uint32_t k = 0; k < EZK_TAB_SIZE; k++) {
            h[k] = acc;
            acc = acc * base;
        }
        Fp big = acc;  // base^(2^14)
        acc = Fp(1);
        for (uint32_t k = 0; k < EZK_TAB_SIZE; k++) {
            h[EZK_TAB_SIZE + k] = acc;
            acc = acc * big;
        }
        return upload(h);
    };
    // compact per-size tables: entry (1 << k) + e = w_{2^k}^e
    auto build_compact = [&](bool inv) {
        std::vector<Fp> h((size_t)2 << kTwMaxLog);
        for (uint32_t k = 0; k <= kTwMaxLog; k++) {
            Fp w = root_of_unity(k);
            if (inv) w = inverse(w);
            Fp acc(1);
            for (uint32_t e = 0; e < (1u << k); e++) {
                h[(1u << k) + e] = acc;
                acc = acc * w;
            }
        }
        return h;
    };
    Fp w = root_of_unity(EZK_ROOT_LOG), o = Fp::from_u64(kDomainOffset);
    t.root_fwd = build(w);
    t.root_inv = build(inverse(w));
    t.off_fwd = build(o);
    t.off_inv = build(inverse(o));
    std::vector<Fp> cf = build_compact(false), ci = build_compact(true);
    const Fp two32 = Fp::from_u64(1ull << 32);
    t.tw_fwd = upload(cf);
    t.tw_inv = upload(ci);
#if EZK_NTT_PRE_TWIDDLES
    auto expand = [&](const std::vector<Fp>& h) {  // entry j -> w_j * 2^(32 i), i = 0..3
        std::vector<Fp> p(4 * h.size());
        for (size_t j = 0; j < h.size(); j++) {
            p[4 * j] = h[j];
            for (int i = 1; i < 4; i++) p[4 * j + i] = p[4 * j + i - 1] * two32;
        }
        return p;
    };
    t.twp_fwd = upload(expand(cf));
    t.twp_inv = upload(expand(ci));
#endif
    Fp w8[2][4][4];
    for (int e = 0; e < 4; e++) {
        w8[0][e][0] = cf[8 + e], w8[1][e][0] = ci[8 + e];
        for (int i = 1; i < 4; i++) w8[0][e][i] = w8[0][e][i - 1] * two32, w8[1][e][i] = w8[1][e][i - 1] * two32;
    }
    EZK_CUDA(cudaMemcpyToSymbol(c_w8pre, w8, sizeof(w8)));
    t.big_tables = new std::map<uint64_t, uint4*>();
    t.scale_table = new NttTables::ScaleTable();
    if (const char* e = getenv("EZK_NTT_BIG_TABLE_MB")) t.big_table_limit_bytes = (uint64_t)atoll(e) << 20;
    const char* env = getenv("EZK_NTT_TILE_LOG");
    if (env) {
        int v = atoi(env);
        if (v >= 6 && v <= 10) t.max_tile_log = v;
    }
}

void ntt_tables_free(NttTables& t) {
    cudaFree(t.root_fwd), cudaFree(t.root_inv), cudaFree(t.off_fwd), cudaFree(t.off_inv);
    cudaFree(t.tw_fwd), cudaFree(t.tw_inv);
    cudaFree(t.twp_fwd), cudaFree(t.twp_inv);
    if (t.big_tables) {
        for (auto& kv : *t.big_tables) cudaFree(kv.second);
        delete t.big_tables;
    }
    if (t.scale_table) {
        cudaFree(t.scale_table->d);
        delete t.scale_table;
    }
    t = NttTables{};
}

int ntt_columns(const NttTables& t, cudaStream_t s, const uint4* src, uint64_t src_pitch, uint4* dst, uint64_t dst_pitch,
                uint4* work, uint32_t ncols, uint32_t log_n, bool inverse, const NttScale* scale) {
    ensure_smem_attr();
    const uint4* roots = inverse ? t.root_inv : t.root_fwd;
    NttScale sc{};
    const uint4* off_tab = t.off_fwd;
    if (scale) {
        sc = *scale;
        sc.enabled = 1;
        off_tab = scale->use_offset == 2 ? t.off_inv : t.off_fwd;
    }
    if (log_n < 3) {
        LaunchScope ls(s, K_NTT_FINAL, ((uint64_t)ncols << log_n) * 32);
        ntt_tiny_kernel<<<ncols, 32, 0, s>>>(src, src_pitch, dst, dst_pitch, log_n, roots, off_tab, sc);
        EZK_CUDA(cudaGetLastError());
        return 1;
    }
    Plan pl = make_plan(log_n, t.max_tile_log);
    const uint64_t n = 1ull << log_n;
    int launches = 0;
    FinalArgs a{};
    if (pl.passes >= 2) {
        launches = run_strided(t, s, pl, log_n, inverse, src, src_pitch, nullptr, work, n, ncols, 0);
        a.src = work, a.src_pitch = n;
    } else {
        a.src = src, a.src_pitch = src_pitch;
    }
    a.dst = dst, a.dst_pitch = dst_pitch;
    a.log_n = log_n, a.log_s = pl.log_d[0], a.passes = pl.passes;
    a.log_top = pl.passes >= 2 ? pl.log_d[pl.passes - 1] : 0;
    a.mode = 0;
    a.lanes_log = pl.passes >= 2 ? lanes_log_for(a.log_s) : 0;
    if (a.lanes_log > a.log_top) a.lanes_log = a.log_top;
    a.inv = inverse ? 1 : 0;
    a.roots = roots;
    a.tw = FinalPass::kPreTw ? (inverse ? t.twp_inv : t.twp_fwd) : (inverse ? t.tw_inv : t.tw_fwd);
    a.off_tab = off_tab;
    a.scale = sc;
    a.scale_tab = scale ? scale_table(t, s, log_n, sc) : nullptr;
    uint64_t runs = 1ull << (log_n - pl.log_d[0]);
    dim3 grid((unsigned)(pl.passes >= 2 ? runs >> a.lanes_log : 1), ncols);
    a.order_log = final_order_log(grid);
    {
        LaunchScope ls(s, K_NTT_FINAL, ((uint64_t)ncols << log_n) * 32);
        launch_final(grid, tile_bytes(a.log_s, a.lanes_log), s, a);
    }
    EZK_CUDA(cudaGetLastError());
    return launches + 1;
}

int composition_recombine(const NttTables& t, cudaStream_t s, const uint4* y, uint32_t log_n, uint64_t m0, uint64_t count,
                          const uint64_t scale[8][2], uint4* out, uint32_t* flag) {
    RecombineConsts k;
    for (int j = 0; j < 8; j++) k.s[j][0] = scale[j][0], k.s[j][1] = scale[j][1];
    {
        LaunchScope ls(s, K_NTT_FINAL, count * 16 * 15);
        composition_recombine_kernel<<<(unsigned)((count + 127) / 128), 128, 0, s>>>(y, log_n, m0, count, t.root_inv, k, out, flag);
    }
    EZK_CUDA(cudaGetLastError());
    return 1;
}

int lde_columns(const NttTables& t, cudaStream_t s, const uint4* coeff, uint64_t coeff_pitch, uint4* lde,
                uint64_t lde_pitch, uint4* tmp, uint32_t ncols, uint32_t log_n, CosetSet cs) {
    ensure_smem_attr();
    if (log_n < 3) throw CudaError("lde_columns: n must be at least 8");
    if (cs.count_log > 3 || cs.base + cs.step * ((1u << cs.count_log) - 1) > 7) throw CudaError("lde_columns: bad coset set");
    Plan pl = make_plan(log_n, t.max_tile_log);
    const uint64_t n = 1ull << log_n;
    const uint32_t log_L = log_n + 3;
    int launches = 0;
    FinalArgs a{};
    a.dst = lde, a.dst_pitch = lde_pitch;
    a.log_n = log_n, a.log_s = pl.log_d[0], a.passes = pl.passes;
    a.log_top = pl.passes >= 2 ? pl.log_d[pl.passes - 1] : 0;
    // lanes: this prover's cosets first, then adjacent runs until the tile is full
    const uint32_t lanes_want = lanes_log_for(a.log_s);
    a.lanes_c_log = lanes_want > cs.count_log ? cs.count_log : lanes_want;
    uint32_t lanes_j_log = pl.passes >= 2 ? lanes_want - a.lanes_c_log : 0;
    if (lanes_j_log > a.log_top) lanes_j_log = a.log_top;
    a.lanes_log = a.lanes_c_log + lanes_j_log;
    a.cs_log = cs.count_log, a.cs_base = cs.base, a.cs_step = cs.step;
    a.log_L = log_L;
    a.inv = 0;
    a.roots = t.root_fwd;
    a.tw = FinalPass::kPreTw ? t.twp_fwd : t.tw_fwd;
    a.off_tab = t.off_fwd;
    if (pl.passes == 1) {
        a.mode = 2;
        a.src = coeff, a.src_pitch = coeff_pitch;
    } else {
        launches += run_strided(t, s, pl, log_n, false, coeff, coeff_pitch, &cs, tmp, n, ncols << cs.count_log, log_L);
        a.mode = 1;
        a.src = tmp, a.src_pitch = n;
    }
    dim3 grid((unsigned)(((1ull << (log_n - pl.log_d[0])) >> lanes_j_log) << (cs.count_log - a.lanes_c_log)), ncols);
    a.order_log = final_order_log(grid);
    {
        const uint64_t elems = (uint64_t)ncols << log_n, outs = elems << cs.count_log;
        LaunchScope ls(s, K_NTT_FINAL, (pl.passes == 1 ? elems + outs : 2 * outs) * 16);
        launch_final(grid, tile_bytes(a.log_s, a.lanes_log), s, a);
    }
    EZK_CUDA(cudaGetLastError());
    return launches + 1;
}

}  // namespace ezk
