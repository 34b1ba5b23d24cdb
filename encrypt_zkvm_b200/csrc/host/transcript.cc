// Host transcript: BLAKE3, random coin, query/proof index logic.  See transcript.h.
#include "transcript.h"
#include <algorithm>
#include <map>
#include <set>
#include <stdexcept>

namespace ezk {

namespace {

const uint32_t kIv[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au, 0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
// message word order for rounds 0..6
const uint8_t kSchedule[7][16] = {
    {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15}, {2, 6, 3, 10, 7, 0, 4, 13, 1, 11, 12, 5, 9, 14, 15, 8},
    {3, 4, 10, 12, 13, 2, 7, 14, 6, 5, 9, 0, 11, 15, 8, 1}, {10, 7, 12, 9, 14, 3, 13, 15, 4, 0, 11, 2, 5, 8, 1, 6},
    {12, 13, 9, 11, 15, 10, 14, 8, 7, 2, 5, 3, 0, 1, 6, 4}, {9, 14, 11, 5, 8, 12, 15, 1, 13, 3, 0, 10, 2, 6, 4, 7},
    {11, 15, 5, 0, 1, 9, 8, 6, 14, 10, 2, 12, 3, 4, 7, 13}};
enum : uint32_t { kChunkStart = 1, kChunkEnd = 2, kParent = 4, kRoot = 8 };

inline uint32_t ror(uint32_t x, unsigned n) { return (x >> n) | (x << (32 - n)); }

struct Cv {
    uint32_t w[8];
};

Cv compress(const Cv& cv, const uint8_t block[64], uint64_t counter, uint32_t len, uint32_t flags) {
    uint32_t m[16];
    memcpy(m, block, 64);
    uint32_t v[16];
    for (int i = 0; i < 8; i++) v[i] = cv.w[i];
    for (int i = 0; i < 4; i++) v[8 + i] = kIv[i];
    v[12] = (uint32_t)counter, v[13] = (uint32_t)(counter >> 32), v[14] = len, v[15] = flags;
    auto mix = [&](int a, int b, int c, int d, uint32_t x, uint32_t y) {
        v[a] += v[b] + x, v[d] = ror(v[d] ^ v[a], 16);
        v[c] += v[d], v[b] = ror(v[b] ^ v[c], 12);
        v[a] += v[b] + y, v[d] = ror(v[d] ^ v[a], 8);
        v[c] += v[d], v[b] = ror(v[b] ^ v[c], 7);
    };
    for (int r = 0; r < 7; r++) {
        const uint8_t* s = kSchedule[r];
        mix(0, 4, 8, 12, m[s[0]], m[s[1]]), mix(1, 5, 9, 13, m[s[2]], m[s[3]]);
        mix(2, 6, 10, 14, m[s[4]], m[s[5]]), mix(3, 7, 11, 15, m[s[6]], m[s[7]]);
        mix(0, 5, 10, 15, m[s[8]], m[s[9]]), mix(1, 6, 11, 12, m[s[10]], m[s[11]]);
        mix(2, 7, 8, 13, m[s[12]], m[s[13]]), mix(3, 4, 9, 14, m[s[14]], m[s[15]]);
    }
    Cv out;
    for (int i = 0; i < 8; i++) out.w[i] = v[i] ^ v[i + 8];
    return out;
}

Cv iv_cv() {
    Cv c;
    memcpy(c.w, kIv, 32);
    return c;
}

Cv chunk_cv(const uint8_t* data, size_t len, uint64_t index, bool is_root) {
    Cv cv = iv_cv();
    size_t nblocks = len ? (len + 63) / 64 : 1;
    for (size_t b = 0; b < nblocks; b++) {
        uint8_t block[64] = {0};
        size_t take = std::min<size_t>(64, len - b * 64);
        memcpy(block, data + b * 64, take);
        uint32_t flags = (b == 0 ? kChunkStart : 0) | (b + 1 == nblocks ? (kChunkEnd | (is_root ? kRoot : 0)) : 0);
        cv = compress(cv, block, index, (uint32_t)take, flags);
    }
    return cv;
}

Cv parent(const Cv& l, const Cv& r, bool is_root) {
    uint8_t block[64];
    memcpy(block, l.w, 32);
    memcpy(block + 32, r.w, 32);
    return compress(iv_cv(), block, 0, 64, kParent | (is_root ? kRoot : 0));
}

}  // namespace

Hash32 blake3(const uint8_t* data, size_t len) {
    Hash32 out;
    const size_t nchunks = len ? (len + 1023) / 1024 : 1;
    if (nchunks == 1) {
        Cv cv = chunk_cv(data, len, 0, true);
        memcpy(out.data(), cv.w, 32);
        return out;
    }
    // chunk-stack merge: after pushing chunk i (0-based), merge while the number of chunks so far has a
    // trailing zero bit; the last chunk is kept aside so the final merges can carry the ROOT flag.
    std::vector<Cv> stack;
    for (size_t i = 0; i + 1 < nchunks; i++) {
        Cv cv = chunk_cv(data + i * 1024, 1024, i, false);
        size_t total = i + 1;
        while ((total & 1) == 0) {
            cv = parent(stack.back(), cv, false);
            stack.pop_back();
            total >>= 1;
        }
        stack.push_back(cv);
    }
    Cv cv = chunk_cv(data + (nchunks - 1) * 1024, len - (nchunks - 1) * 1024, nchunks - 1, false);
    while (!stack.empty()) {
        cv = parent(stack.back(), cv, stack.size() == 1);
        stack.pop_back();
    }
    memcpy(out.data(), cv.w, 32);
    return out;
}

Hash32 hash_elements(const Fp* e, size_t n) { return blake3(reinterpret_cast<const uint8_t*>(e), n * 16); }

Hash32 merge_digests(const Hash32& a, const Hash32& b) {
    uint8_t buf[64];
    memcpy(buf, a.data(), 32);
    memcpy(buf + 32, b.data(), 32);
    return blake3(buf, 64);
}

Hash32 merge_with_int(const Hash32& seed, uint64_t v) {
    uint8_t buf[40];
    memcpy(buf, seed.data(), 32);
    memcpy(buf + 32, &v, 8);
    return blake3(buf, 40);
}

void RandomCoin::init(const std::vector<Fp>& seed_elements) {
    seed_ = hash_elements(seed_elements.data(), seed_elements.size());
    counter_ = 0;
}
void RandomCoin::reseed(const Hash32& d) {
    seed_ = merge_digests(seed_, d);
    counter_ = 0;
}
Hash32 RandomCoin::next() {
    counter_ += 1;
    return merge_with_int(seed_, counter_);
}
Fp RandomCoin::draw() {
    for (int attempt = 0; attempt < 1000; attempt++) {
        Hash32 h = next();
        Fp v = fp_load(h.data());
        if (v.v < Fp::modulus()) return v;
    }
    throw std::runtime_error("random coin: no valid field element in 1000 draws");
}
unsigned RandomCoin::leading_zeros(uint64_t nonce) const {
    Hash32 h = merge_with_int(seed_, nonce);
    uint64_t head;
    memcpy(&head, h.data(), 8);
    return head ? (unsigned)__builtin_ctzll(head) : 64;
}
std::vector<uint64_t> RandomCoin::draw_integers(size_t count, uint64_t domain_size, uint64_t nonce) {
    seed_ = merge_with_int(seed_, nonce);
    counter_ = 0;
    std::vector<uint64_t> out(count);
    for (auto& v : out) {
        Hash32 h = next();
        uint64_t head;
        memcpy(&head, h.data(), 8);
        v = head & (domain_size - 1);
    }
    return out;
}

WireCompat& wire_compat() {
    static WireCompat c;
    return c;
}

size_t num_fri_layers(uint64_t lde_size, const ProofOptions& o) {
    const uint64_t max_remainder = (uint64_t)(o.fri_rem_max_deg + 1) * o.blowup;
    size_t layers = 0;
    for (; lde_size > max_remainder; lde_size /= o.fri_fold) layers++;
    return layers;
}

std::vector<Fp> coin_seed(uint32_t trace_width, uint64_t trace_len, const ProofOptions& o, const Fp pub_inputs[18]) {
    std::vector<Fp> e;
    e.reserve(26);
    e.push_back(Fp::from_u64(((uint64_t)trace_width << 8) | 0));  // main width, no auxiliary segments
    e.push_back(Fp::from_u64((uint32_t)trace_len));
    const u128 m = Fp::modulus();
    e.push_back(Fp::from_u64((uint64_t)m));          // modulus bytes 0..8
    e.push_back(Fp::from_u64((uint64_t)(m >> 64)));  // modulus bytes 8..16
    e.push_back(Fp::from_u64(((uint64_t)o.field_ext << 16) | ((uint64_t)o.fri_fold << 8) | o.fri_rem_max_deg));
    e.push_back(Fp::from_u64(o.grinding));
    e.push_back(Fp::from_u64(o.blowup));
    e.push_back(Fp::from_u64(o.num_queries));
    for (int i = 0; i < 18; i++) e.push_back(pub_inputs[i]);
    return e;
}

std::vector<uint64_t> fold_positions(const std::vector<uint64_t>& positions, uint64_t domain_size, uint64_t folding) {
    const uint64_t target = domain_size / folding;
    std::vector<uint64_t> out;
    for (uint64_t p : positions) {
        uint64_t q = p % target;
        if (std::find(out.begin(), out.end(), q) == out.end()) out.push_back(q);
    }
    return out;
}

std::vector<std::vector<uint64_t>> batch_proof_node_indices(uint64_t num_leaves, const std::vector<uint64_t>& leaf_indexes) {
    std::set<uint64_t> queried(leaf_indexes.begin(), leaf_indexes.end());
    if (queried.size() != leaf_indexes.size()) throw std::runtime_error("batch proof: duplicate leaf index");
    std::set<uint64_t> pairs;
    for (uint64_t i : leaf_indexes) pairs.insert(i & ~1ull);
    std::vector<std::vector<uint64_t>> nodes;
    std::vector<uint64_t> level;
    for (uint64_t first : pairs) {
        std::vector<uint64_t> v;
        for (uint64_t i = first; i < first + 2; i++)
            if (!queried.count(i)) v.push_back(num_leaves + i);
        nodes.push_back(v);
        level.push_back((first + num_leaves) >> 1);
    }
    unsigned depth = 0;
    while ((1ull << depth) < num_leaves) depth++;
    for (unsigned d = 1; d < depth; d++) {
        std::vector<uint64_t> parents;
        for (size_t i = 0; i < level.size(); i++) {
            const uint64_t sibling = level[i] ^ 1;
            if (i + 1 < level.size() && level[i + 1] == sibling) {
                i++;  // both children are known: nothing to send
            } else {
                // Winterfell files the sibling under the *current-level position* of this path
                nodes[i].push_back(sibling);
            }
            parents.push_back(sibling >> 1);
        }
        level.swap(parents);
    }
    return nodes;
}

void ProofWriter::u16(uint16_t v) {
    for (int i = 0; i < 2; i++) buf_.push_back((uint8_t)(v >> (8 * i)));
}
void ProofWriter::u32(uint32_t v) {
    for (int i = 0; i < 4; i++) buf_.push_back((uint8_t)(v >> (8 * i)));
}
void ProofWriter::u64(uint64_t v) {
    for (int i = 0; i < 8; i++) buf_.push_back((uint8_t)(v >> (8 * i)));
}
void ProofWriter::bytes(const void* p, size_t n) {
    const uint8_t* b = static_cast<const uint8_t*>(p);
    buf_.insert(buf_.end(), b, b + n);
}
void ProofWriter::element(Fp v) {
    uint8_t t[16];
    fp_store(t, v);
    bytes(t, 16);
}

}  // namespace ezk
