// Device BLAKE3-256 compression for the prover's hasher `Blake3_256<BaseElement>`
// (prover/src/lib.rs:13,44).  Every input on the proving path is a single chunk (<= 1024 bytes):
// rows of 28 / 7 / 8 field elements (448 / 112 / 128 bytes) and 64-byte node merges, so only the
// chunk-chaining part of BLAKE3 is needed on the GPU.  State lives in registers; the message
// schedule is resolved at compile time (fully unrolled rounds, constant indices).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace ezk {
namespace dev {

#define B3_CHUNK_START 1u
#define B3_CHUNK_END 2u
#define B3_PARENT 4u
#define B3_ROOT 8u

#define B3_IV0 0x6A09E667u
#define B3_IV1 0xBB67AE85u
#define B3_IV2 0x3C6EF372u
#define B3_IV3 0xA54FF53Au
#define B3_IV4 0x510E527Fu
#define B3_IV5 0x9B05688Cu
#define B3_IV6 0x1F83D9ABu
#define B3_IV7 0x5BE0CD19u

__device__ __forceinline__ uint32_t b3_rotr(uint32_t x, int n) { return __funnelshift_r(x, x, n); }

// Measured and kept out: issuing the additions as integer multiply-adds (x * 1 + y with the 1 read from constant
// memory, so that they run on the FMA pipe: 8 ALU + 6 FMA instructions per G instead of 10 + 2) leaves the hash
// kernels' time unchanged on the B200 (2.70 -> 2.68 ms per 2^20 proof) although the ALU pipe reads 96 % busy.
#define B3_G(a, b, c, d, mx, my)   \
    a = a + b + (mx);              \
    d = __byte_perm(d ^ a, 0, 0x1032); /* rotr 16 */ \
    c = c + d;                     \
    b = b3_rotr(b ^ c, 12);        \
    a = a + b + (my);              \
    d = __byte_perm(d ^ a, 0, 0x0321); /* rotr 8 */  \
    c = c + d;                     \
    b = b3_rotr(b ^ c, 7);

#define B3_ROUND(m, i0, i1, i2, i3, i4, i5, i6, i7, i8, i9, i10, i11, i12, i13, i14, i15) \
    B3_G(s0, s4, s8, s12, m[i0], m[i1])                                                    \
    B3_G(s1, s5, s9, s13, m[i2], m[i3])                                                    \
    B3_G(s2, s6, s10, s14, m[i4], m[i5])                                                   \
    B3_G(s3, s7, s11, s15, m[i6], m[i7])                                                   \
    B3_G(s0, s5, s10, s15, m[i8], m[i9])                                                   \
    B3_G(s1, s6, s11, s12, m[i10], m[i11])                                                 \
    B3_G(s2, s7, s8, s13, m[i12], m[i13])                                                  \
    B3_G(s3, s4, s9, s14, m[i14], m[i15])

// cv[8] <- compress(cv, m[16], counter = 0, block_len, flags), truncated to the chaining value
__device__ __forceinline__ void b3_compress(uint32_t cv[8], const uint32_t m[16], uint32_t block_len, uint32_t flags) {
    uint32_t s0 = cv[0], s1 = cv[1], s2 = cv[2], s3 = cv[3], s4 = cv[4], s5 = cv[5], s6 = cv[6], s7 = cv[7];
    uint32_t s8 = B3_IV0, s9 = B3_IV1, s10 = B3_IV2, s11 = B3_IV3, s12 = 0, s13 = 0, s14 = block_len, s15 = flags;
    B3_ROUND(m, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15)
    B3_ROUND(m, 2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8)
    B3_ROUND(m, 3, 4, 10, 12, 13, 2, 7, 14, 6, 5, 9, 0, 11, 15, 8, 1)
    B3_ROUND(m, 10, 7, 12, 9, 14, 3, 13, 15, 4, 0, 11, 2, 5, 8, 1, 6)
    B3_ROUND(m, 12, 13, 9, 11, 15, 10, 14, 8, 7, 2, 5, 3, 0, 1, 6, 4)
    B3_ROUND(m, 9, 14, 11, 5, 8, 12, 15, 1, 13, 3, 0, 10, 2, 6, 4, 7)
    B3_ROUND(m, 11, 15, 5, 0, 1, 9, 8, 6, 14, 10, 2, 12, 3, 4, 7, 13)
    cv[0] = s0 ^ s8, cv[1] = s1 ^ s9, cv[2] = s2 ^ s10, cv[3] = s3 ^ s11;
    cv[4] = s4 ^ s12, cv[5] = s5 ^ s13, cv[6] = s6 ^ s14, cv[7] = s7 ^ s15;
}

__device__ __forceinline__ void b3_init(uint32_t cv[8]) {
    cv[0] = B3_IV0, cv[1] = B3_IV1, cv[2] = B3_IV2, cv[3] = B3_IV3;
    cv[4] = B3_IV4, cv[5] = B3_IV5, cv[6] = B3_IV6, cv[7] = B3_IV7;
}

// merge(a, b) = blake3(a || b): one 64-byte block, flags START|END|ROOT
__device__ __forceinline__ void b3_merge(const uint4* left_right /* 4 x uint4 = 64 bytes */, uint4 out[2]) {
    uint32_t m[16];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        uint4 v = left_right[i];
        m[4 * i] = v.x, m[4 * i + 1] = v.y, m[4 * i + 2] = v.z, m[4 * i + 3] = v.w;
    }
    uint32_t cv[8];
    b3_init(cv);
    b3_compress(cv, m, 64, B3_CHUNK_START | B3_CHUNK_END | B3_ROOT);
    out[0] = make_uint4(cv[0], cv[1], cv[2], cv[3]);
    out[1] = make_uint4(cv[4], cv[5], cv[6], cv[7]);
}

}  // namespace dev
}  // namespace ezk
