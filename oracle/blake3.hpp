// ORACLE (test infrastructure, not product): portable BLAKE3-256 (default hash mode, any length).
//
// The reference configures `HashFn = Blake3_256<BaseElement>` (prover/src/lib.rs:13,44) which
// calls the `blake3` crate 1.5.4 (Cargo.lock:47-58).  This restates the published BLAKE3
// algorithm (7 rounds, 1024-byte chunks, binary tree of parent nodes); it is pinned against
// the in-container `blake3` Python module and the official test-vector pattern in
// tests/test_oracle_blake3.py + tests/golden/blake3_kat.json.
#pragma once
#include <cstdint>
#include <cstring>
#include <cstddef>

namespace orc {
namespace b3 {

static const uint32_t IV[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au,
                               0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
static const uint8_t PERM[16] = {2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8};
enum : uint32_t { CHUNK_START = 1, CHUNK_END = 2, PARENT = 4, ROOT = 8 };

static inline uint32_t rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }

static inline void g(uint32_t* s, int a, int b, int c, int d, uint32_t mx, uint32_t my) {
    s[a] = s[a] + s[b] + mx;
    s[d] = rotr(s[d] ^ s[a], 16);
    s[c] = s[c] + s[d];
    s[b] = rotr(s[b] ^ s[c], 12);
    s[a] = s[a] + s[b] + my;
    s[d] = rotr(s[d] ^ s[a], 8);
    s[c] = s[c] + s[d];
    s[b] = rotr(s[b] ^ s[c], 7);
}

// out[16] = full compression output; out[0..8] is the chaining value
static inline void compress(const uint32_t cv[8], const uint32_t block[16], uint64_t counter,
                            uint32_t block_len, uint32_t flags, uint32_t out[16]) {
    uint32_t s[16] = {cv[0], cv[1], cv[2], cv[3], cv[4], cv[5], cv[6], cv[7],
                      IV[0], IV[1], IV[2], IV[3], (uint32_t)counter, (uint32_t)(counter >> 32),
                      block_len, flags};
    uint32_t m[16];
    memcpy(m, block, 64);
    for (int r = 0; r < 7; r++) {
        g(s, 0, 4, 8, 12, m[0], m[1]);
        g(s, 1, 5, 9, 13, m[2], m[3]);
        g(s, 2, 6, 10, 14, m[4], m[5]);
        g(s, 3, 7, 11, 15, m[6], m[7]);
        g(s, 0, 5, 10, 15, m[8], m[9]);
        g(s, 1, 6, 11, 12, m[10], m[11]);
        g(s, 2, 7, 8, 13, m[12], m[13]);
        g(s, 3, 4, 9, 14, m[14], m[15]);
        uint32_t p[16];
        for (int i = 0; i < 16; i++) p[i] = m[PERM[i]];
        memcpy(m, p, 64);
    }
    for (int i = 0; i < 8; i++) {
        out[i] = s[i] ^ s[i + 8];
        out[i + 8] = s[i + 8] ^ cv[i];
    }
}

// chaining value of one chunk (<= 1024 bytes); `root` only when the whole input is this chunk
static inline void chunk_cv(const uint8_t* data, size_t len, uint64_t chunk_index, bool root,
                            uint32_t cv_out[8]) {
    uint32_t cv[8];
    memcpy(cv, IV, 32);
    size_t nblocks = len == 0 ? 1 : (len + 63) / 64;
    uint32_t out[16];
    for (size_t b = 0; b < nblocks; b++) {
        uint32_t block[16] = {0};
        size_t off = b * 64;
        size_t blen = len - off < 64 ? len - off : 64;
        memcpy(block, data + off, blen);  // little-endian host
        uint32_t flags = 0;
        if (b == 0) flags |= CHUNK_START;
        if (b == nblocks - 1) flags |= CHUNK_END | (root ? (uint32_t)ROOT : 0u);
        compress(cv, block, chunk_index, (uint32_t)blen, flags, out);
        memcpy(cv, out, 32);
    }
    memcpy(cv_out, cv, 32);
}

static inline void parent_cv(const uint32_t l[8], const uint32_t r[8], bool root, uint32_t out_cv[8]) {
    uint32_t block[16], out[16];
    memcpy(block, l, 32);
    memcpy(block + 8, r, 32);
    compress(IV, block, 0, 64, PARENT | (root ? (uint32_t)ROOT : 0u), out);
    memcpy(out_cv, out, 32);
}

// recursive tree: left subtree takes the largest power-of-two number of chunks < total
static inline void subtree(const uint8_t* data, size_t len, uint64_t chunk0, bool root, uint32_t cv[8]) {
    if (len <= 1024) {
        chunk_cv(data, len, chunk0, root, cv);
        return;
    }
    size_t chunks = (len + 1023) / 1024;
    size_t left = 1;
    while (left * 2 < chunks) left *= 2;
    uint32_t l[8], r[8];
    subtree(data, left * 1024, chunk0, false, l);
    subtree(data + left * 1024, len - left * 1024, chunk0 + left, false, r);
    parent_cv(l, r, root, cv);
}

static inline void hash(const uint8_t* data, size_t len, uint8_t out[32]) {
    uint32_t cv[8];
    subtree(data, len, 0, true, cv);
    memcpy(out, cv, 32);
}

}  // namespace b3

struct Digest {
    uint8_t b[32];
    bool operator==(const Digest& o) const { return memcmp(b, o.b, 32) == 0; }
    bool operator!=(const Digest& o) const { return !(*this == o); }
};

// winter-crypto 0.9.0 `Blake3_256` (restated)
static inline Digest hash_bytes(const uint8_t* p, size_t n) {
    Digest d;
    b3::hash(p, n, d.b);
    return d;
}
static inline Digest merge(const Digest& a, const Digest& b) {
    uint8_t buf[64];
    memcpy(buf, a.b, 32);
    memcpy(buf + 32, b.b, 32);
    return hash_bytes(buf, 64);
}
static inline Digest merge_with_int(const Digest& seed, uint64_t v) {
    uint8_t buf[40];
    memcpy(buf, seed.b, 32);
    memcpy(buf + 32, &v, 8);
    return hash_bytes(buf, 40);
}

}  // namespace orc
