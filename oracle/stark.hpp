// ORACLE (test infrastructure, not product): single-threaded CPU restatement of the reference's
// prove + verify path.
//
// PARITY STATUS: the engine behind `ExecutionProver::prove` (prover/src/lib.rs:40-77, called from
// vm/src/lib.rs:26) is the un-vendored crate winterfell 0.9.0 (Cargo.toml:13, Cargo.lock:537-636).
// No Rust toolchain and no copy of that crate exist in this environment, and the reference's own
// tests hold NO byte-level golden vectors for this path => "parity unpinned" for the
// transcript / serialization layer: this file restates the published algorithm,
// byte-level guesses are isolated in `Compat`.  What IS pinned: the field and BLAKE3 (Python
// big-int / `blake3` module), the AIR (reference's own 15 unit-test frames, tests/test_air_frames.py),
// and the algebra (the restated verifier accepts, 1-bit mutations are rejected).
#pragma once
#include "air.hpp"
#include "blake3.hpp"
#include "f128.hpp"
#include "ntt.hpp"
#include <algorithm>
#include <map>
#include <set>
#include <stdexcept>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace orc {

// Byte-level details that could not be checked against the real crate.
struct Compat {
    bool ood_interleaved = true;       // A.7: [cur_0, next_0, cur_1, next_1, ...]
    bool remainder_low_to_high = true; // A.9
    uint64_t first_nonce = 1;          // A.3 (7): grind search starts at 1
    // A.10: TraceInfo::write_into stores main width, aux width AND the aux segment's random-element count
    // (all u8) before log2(length): TraceInfo::read_from rebuilds itself through new_multi_segment(main, aux,
    // num_aux_rands, length, meta), so the count has to be on the wire.  false = the two-byte form.
    bool trace_info_aux_rands_byte = true;
};

struct ProofOptions {  // vm/src/lib.rs:20
    uint32_t num_queries = 32, blowup = 8, grinding = 0, field_ext = 1, fri_fold = 8, fri_rem_max_deg = 127;
};

static inline Digest hash_elements(const u128* e, size_t n) { return hash_bytes((const uint8_t*)e, n * 16); }

// winter-crypto DefaultRandomCoin<Blake3_256>
struct Coin {
    Digest seed;
    uint64_t counter = 0;
    void init(const std::vector<u128>& elems) {
        seed = hash_elements(elems.data(), elems.size());
        counter = 0;
    }
    void reseed(const Digest& d) {
        seed = merge(seed, d);
        counter = 0;
    }
    Digest next() {
        counter++;
        return merge_with_int(seed, counter);
    }
    u128 draw() {
        for (int i = 0; i < 1000; i++) {
            Digest d = next();
            u128 v = load_le(d.b);
            if (v < MOD) return v;
        }
        throw std::runtime_error("coin: failed to draw element");
    }
    unsigned leading_zeros(uint64_t v) const {
        Digest d = merge_with_int(seed, v);
        uint64_t head;
        memcpy(&head, d.b, 8);
        return head == 0 ? 64 : (unsigned)__builtin_ctzll(head);
    }
    std::vector<size_t> draw_integers(size_t num, size_t domain, uint64_t nonce) {
        seed = merge_with_int(seed, nonce);
        counter = 0;
        std::vector<size_t> out;
        for (size_t i = 0; i < num; i++) {
            Digest d = next();
            uint64_t head;
            memcpy(&head, d.b, 8);
            out.push_back((size_t)(head & (uint64_t)(domain - 1)));
        }
        return out;
    }
};

// winter-crypto MerkleTree: nodes[1] root, nodes[k] = merge(nodes[2k], nodes[2k+1])
struct MerkleTree {
    std::vector<Digest> nodes;  // size 2*leaves; [leaves..2*leaves) are the leaves
    size_t num_leaves = 0;
    unsigned depth = 0;
    void build(std::vector<Digest>&& leaves) {
        num_leaves = leaves.size();
        depth = ilog2(num_leaves);
        nodes.resize(2 * num_leaves);
        for (size_t i = 0; i < num_leaves; i++) nodes[num_leaves + i] = leaves[i];
        for (size_t lvl = num_leaves / 2; lvl >= 1; lvl >>= 1) {
#pragma omp parallel for schedule(static) if (lvl >= 4096)
            for (size_t k = lvl; k < 2 * lvl; k++) nodes[k] = merge(nodes[2 * k], nodes[2 * k + 1]);
        }
    }
    const Digest& root() const { return nodes[1]; }
    const Digest& leaf(size_t i) const { return nodes[num_leaves + i]; }
};

struct BatchProof {
    std::vector<Digest> leaves;               // in input order
    std::vector<std::vector<Digest>> nodes;   // one vector per normalized index
    unsigned depth = 0;
};

static inline std::vector<size_t> normalize_indexes(const std::vector<size_t>& idx) {
    std::set<size_t> s;
    for (size_t i : idx) s.insert(i & ~(size_t)1);
    return std::vector<size_t>(s.begin(), s.end());
}

// MerkleTree::prove_batch (App. A.10), including the "nodes[i] indexed by current-level position" behaviour
static inline BatchProof prove_batch(const MerkleTree& t, const std::vector<size_t>& indexes) {
    std::map<size_t, size_t> index_map;
    for (size_t i = 0; i < indexes.size(); i++) index_map[indexes[i]] = i;
    if (index_map.size() != indexes.size()) throw std::runtime_error("prove_batch: duplicate index");
    auto norm = normalize_indexes(indexes);
    BatchProof p;
    p.depth = t.depth;
    p.leaves.resize(indexes.size());
    std::vector<size_t> next;
    size_t n = t.num_leaves;
    for (size_t index : norm) {
        std::vector<Digest> missing;
        for (size_t i = index; i < index + 2; i++) {
            auto it = index_map.find(i);
            if (it != index_map.end())
                p.leaves[it->second] = t.leaf(i);
            else
                missing.push_back(t.leaf(i));
        }
        p.nodes.push_back(missing);
        next.push_back((index + n) >> 1);
    }
    for (unsigned d = 1; d < t.depth; d++) {
        std::vector<size_t> cur = next;
        next.clear();
        size_t i = 0;
        while (i < cur.size()) {
            size_t sib = cur[i] ^ 1;
            if (i + 1 < cur.size() && cur[i + 1] == sib)
                i += 1;
            else
                p.nodes[i].push_back(t.nodes[sib]);
            next.push_back(sib >> 1);
            i += 1;
        }
    }
    return p;
}

static inline std::vector<uint8_t> serialize_nodes(const BatchProof& p) {
    std::vector<uint8_t> out;
    out.push_back((uint8_t)p.nodes.size());
    for (auto& v : p.nodes) {
        out.push_back((uint8_t)v.size());
        for (auto& d : v) out.insert(out.end(), d.b, d.b + 32);
    }
    return out;
}

// BatchMerkleProof::get_root (verifier side, App. A.10/A.11). Returns false on malformed proofs.
static inline bool batch_root(const BatchProof& p, const std::vector<size_t>& indexes, Digest& root_out) {
    if (indexes.empty()) return false;
    std::map<size_t, size_t> index_map;
    for (size_t i = 0; i < indexes.size(); i++) {
        if (indexes[i] >= ((size_t)1 << p.depth)) return false;
        index_map[indexes[i]] = i;
    }
    if (index_map.size() != indexes.size()) return false;
    auto norm = normalize_indexes(indexes);
    if (norm.size() != p.nodes.size()) return false;
    std::map<size_t, Digest> v;
    size_t offset = (size_t)1 << p.depth;
    std::vector<size_t> next, ptr;
    for (size_t i = 0; i < norm.size(); i++) {
        size_t index = norm[i];
        Digest buf[2];
        auto i1 = index_map.find(index), i2 = index_map.find(index + 1);
        if (i1 != index_map.end()) {
            if (p.leaves.size() <= i1->second) return false;
            buf[0] = p.leaves[i1->second];
            if (i2 != index_map.end()) {
                if (p.leaves.size() <= i2->second) return false;
                buf[1] = p.leaves[i2->second];
                ptr.push_back(0);
            } else {
                if (p.nodes[i].empty()) return false;
                buf[1] = p.nodes[i][0];
                ptr.push_back(1);
            }
        } else {
            if (p.nodes[i].empty()) return false;
            buf[0] = p.nodes[i][0];
            if (i2 == index_map.end() || p.leaves.size() <= i2->second) return false;
            buf[1] = p.leaves[i2->second];
            ptr.push_back(1);
        }
        size_t parent = (offset + index) >> 1;
        v[parent] = merge(buf[0], buf[1]);
        next.push_back(parent);
    }
    for (unsigned d = 1; d < p.depth; d++) {
        std::vector<size_t> cur = next;
        next.clear();
        size_t i = 0;
        while (i < cur.size()) {
            size_t node_index = cur[i], sib_index = node_index ^ 1;
            Digest sib;
            if (i + 1 < cur.size() && cur[i + 1] == sib_index) {
                auto it = v.find(sib_index);
                if (it == v.end()) return false;
                sib = it->second;
                i += 1;
            } else {
                size_t pointer = ptr[i];
                if (p.nodes[i].size() <= pointer) return false;
                sib = p.nodes[i][pointer];
                ptr[i] += 1;
            }
            auto it = v.find(node_index);
            if (it == v.end()) return false;
            Digest parent = (node_index & 1) ? merge(sib, it->second) : merge(it->second, sib);
            v[node_index >> 1] = parent;
            next.push_back(node_index >> 1);
            i += 1;
        }
    }
    auto it = v.find(1);
    if (it == v.end()) return false;
    root_out = it->second;
    return true;
}

static inline bool parse_nodes(const uint8_t* p, size_t len, std::vector<std::vector<Digest>>& out) {
    size_t pos = 0;
    if (len < 1) return false;
    size_t nv = p[pos++];
    out.clear();
    for (size_t i = 0; i < nv; i++) {
        if (pos >= len) return false;
        size_t nd = p[pos++];
        if (pos + nd * 32 > len) return false;
        std::vector<Digest> v(nd);
        for (size_t k = 0; k < nd; k++) memcpy(v[k].b, p + pos + 32 * k, 32);
        pos += nd * 32;
        out.push_back(v);
    }
    return pos == len;
}

// ---------------------------------------------------------------------------------------------
// domain / context numbers (App. A.2)
static inline size_t num_fri_layers(size_t domain, const ProofOptions& o) {
    size_t max_rem = (size_t)(o.fri_rem_max_deg + 1) * o.blowup, r = 0;
    while (domain > max_rem) {
        domain /= o.fri_fold;
        r++;
    }
    return r;
}

// Context::to_elements ++ PublicInputs::to_elements (App. A.1; air/src/lib.rs:38-47)
static inline std::vector<u128> coin_seed_elements(size_t n, const ProofOptions& o, const u128* pub18) {
    std::vector<u128> e;
    e.push_back(((u128)TRACE_WIDTH << 8) | 0);
    e.push_back((u128)(uint32_t)n);
    e.push_back((u128)(uint64_t)MOD);          // low 8 bytes of the modulus
    e.push_back((u128)(uint64_t)(MOD >> 64));  // high 8 bytes
    e.push_back(((u128)o.field_ext << 16) | ((u128)o.fri_fold << 8) | o.fri_rem_max_deg);
    e.push_back(o.grinding);
    e.push_back(o.blowup);
    e.push_back(o.num_queries);
    for (int i = 0; i < 18; i++) e.push_back(pub18[i]);
    return e;
}

static inline std::vector<size_t> fold_positions(const std::vector<size_t>& pos, size_t domain, size_t fold) {
    size_t target = domain / fold;
    std::vector<size_t> out;
    for (size_t p : pos) {
        size_t q = p % target;
        if (std::find(out.begin(), out.end(), q) == out.end()) out.push_back(q);
    }
    return out;
}

struct ByteWriter {
    std::vector<uint8_t> b;
    void u8(uint8_t v) { b.push_back(v); }
    void u16(uint16_t v) { b.push_back(v & 0xFF), b.push_back(v >> 8); }
    void u32(uint32_t v) {
        for (int i = 0; i < 4; i++) b.push_back((v >> (8 * i)) & 0xFF);
    }
    void u64(uint64_t v) {
        for (int i = 0; i < 8; i++) b.push_back((v >> (8 * i)) & 0xFF);
    }
    void bytes(const uint8_t* p, size_t n) { b.insert(b.end(), p, p + n); }
    void elem(u128 v) {
        uint8_t t[16];
        store_le(t, v);
        bytes(t, 16);
    }
};

// Everything the prover computes on the way, kept for stage-by-stage parity checks.
struct Artifacts {
    size_t n = 0, L = 0;
    std::vector<std::vector<u128>> trace_polys;     // 28 x n coefficients
    std::vector<u128> trace_lde;                    // L x 28 row-major
    Digest trace_root;
    std::vector<u128> tcoef, bcoef;                 // 20 + 22
    std::vector<u128> combined;                     // L
    std::vector<std::vector<u128>> comp_polys;      // 7 x n
    std::vector<u128> comp_lde;                     // L x 7 row-major
    Digest comp_root;
    u128 z = 0;
    std::vector<u128> ood_cur, ood_next, ood_comp;  // 28, 28, 7
    std::vector<u128> deep_tc, deep_cc;             // 28, 7
    std::vector<u128> deep_evals;                   // L
    size_t deep_degree = 0;
    std::vector<Digest> fri_roots;                  // per layer, then remainder commitment
    std::vector<u128> fri_alphas;
    std::vector<std::vector<u128>> fri_layer_evals; // evaluations entering each layer
    std::vector<u128> remainder;
    uint64_t pow_nonce = 0;
    std::vector<size_t> positions;
    std::vector<uint8_t> proof;
};

static inline size_t poly_degree(const std::vector<u128>& p) {
    for (size_t i = p.size(); i-- > 0;)
        if (p[i] != 0) return i;
    return 0;
}

struct ProveError : std::runtime_error {
    int code;
    ProveError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

// winter-prover 0.9.0 `Prover::prove` -> generate_proof, as reached from
// vm/src/lib.rs:26 through prover/src/lib.rs:40-77.
// cols: 28 pointers to n canonical elements (TraceTable columns, vm/src/lib.rs:18).
static inline void prove(const u128* const* cols, size_t n, const u128* pub18, const AirParams& ap,
                         const ProofOptions& opt, const Compat& cp, Artifacts& A) {
    const size_t B = opt.blowup, L = n * B, W = TRACE_WIDTH, C = NUM_COMP_COLUMNS;
    const u128 o = GENERATOR;
    const unsigned lgL = ilog2(L), lgn = ilog2(n);
    A.n = n, A.L = L;
    if (n < 16 || (n & (n - 1))) throw ProveError(1, "trace length must be a power of two >= 16");

    // (0) channel: coin seeded with context + public inputs (App. A.1)
    Coin coin;
    coin.init(coin_seed_elements(n, opt, pub18));
    std::vector<uint8_t> commitments;

    // (1) trace LDE + commitment (App. A.4; prover/src/lib.rs:55-62)
    A.trace_polys.assign(W, std::vector<u128>());
    A.trace_lde.assign(L * W, 0);
#pragma omp parallel for schedule(dynamic)
    for (size_t c = 0; c < W; c++) {
        std::vector<u128> p(cols[c], cols[c] + n);
        interpolate_poly(p);
        auto ev = evaluate_poly_with_offset(p, o, B);
        for (size_t i = 0; i < L; i++) A.trace_lde[i * W + c] = ev[i];
        A.trace_polys[c] = std::move(p);
    }
    MerkleTree trace_tree;
    {
        std::vector<Digest> leaves(L);
#pragma omp parallel for schedule(static)
        for (size_t i = 0; i < L; i++) leaves[i] = hash_elements(&A.trace_lde[i * W], W);
        trace_tree.build(std::move(leaves));
    }
    A.trace_root = trace_tree.root();
    commitments.insert(commitments.end(), A.trace_root.b, A.trace_root.b + 32);
    coin.reseed(A.trace_root);

    // (2) constraint composition coefficients: 20 transition then 22 boundary draws (App. A.3)
    A.tcoef.resize(NUM_TRANSITION), A.bcoef.resize(NUM_ASSERTIONS);
    for (auto& x : A.tcoef) x = coin.draw();
    for (auto& x : A.bcoef) x = coin.draw();

    // evaluate constraints over the LDE domain (App. A.5; prover/src/lib.rs:65-72)
    auto asserts = sorted_assertions(n, pub18);
    const u128 g = root_of_unity(lgn), wL = root_of_unity(lgL);
    const u128 g_last = fexp(g, n - NUM_EXEMPTIONS);       // g^(n-2)
    const u128 g_last2 = fexp(g, n - 1);                   // g^(n-1)
    // periodic table: 128 rows x 9 (row = step mod 128): P_p(x^(n/16)), x = o*w_L^step
    const size_t PT = CYCLE_LENGTH * B;
    std::vector<u128> ptable(PT * NUM_PERIODIC);
    {
        auto polys = periodic_polys();
        u128 on = fexp(o, n / CYCLE_LENGTH), w128 = root_of_unity(ilog2(PT));
        for (size_t r = 0; r < PT; r++) {
            u128 y = fmul(on, fexp(w128, r));
            for (unsigned p = 0; p < NUM_PERIODIC; p++) ptable[r * NUM_PERIODIC + p] = eval_horner(polys[p].data(), CYCLE_LENGTH, y);
        }
    }
    A.combined.assign(L, 0);
    {
        std::vector<u128> tnum(L), b0(L), b1(L), d0(L), d1(L), dz(L);
        auto xs = power_table(wL, L);
        std::vector<u128> zn_minus_1(B);
        {
            u128 on = fexp(o, n), wB = root_of_unity(ilog2(B)), acc = 1;
            for (size_t c = 0; c < B; c++) {
                zn_minus_1[c] = fsub(fmul(on, acc), 1);
                acc = fmul(acc, wB);
            }
        }
#pragma omp parallel for schedule(static)
        for (size_t i = 0; i < L; i++) {
            u128 x = fmul(o, xs[i]);
            const u128* cur = &A.trace_lde[i * W];
            const u128* nxt = &A.trace_lde[((i + B) % L) * W];
            u128 ev[NUM_TRANSITION];
            evaluate_transition(cur, nxt, &ptable[(i % PT) * NUM_PERIODIC], ap, ev);
            u128 t = 0;
            for (unsigned j = 0; j < NUM_TRANSITION; j++) t = fadd(t, fmul(A.tcoef[j], ev[j]));
            // transition divisor (x^n - 1) / ((x - g^(n-2)) (x - g^(n-1)))
            tnum[i] = fmul(t, fmul(fsub(x, g_last), fsub(x, g_last2)));
            dz[i] = zn_minus_1[i % B];  // x^n - 1 takes only B distinct values: x^n = o^n * w_B^(i mod B)
            u128 s0 = 0, s1 = 0;
            for (unsigned k = 0; k < NUM_ASSERTIONS; k++) {
                u128 term = fmul(A.bcoef[k], fsub(cur[asserts[k].column], asserts[k].value));
                if (asserts[k].step == 0)
                    s0 = fadd(s0, term);
                else
                    s1 = fadd(s1, term);
            }
            b0[i] = s0, b1[i] = s1;
            d0[i] = fsub(x, 1);
            d1[i] = fsub(x, g_last);
        }
        batch_inverse(dz.data(), L);
        batch_inverse(d0.data(), L);
        batch_inverse(d1.data(), L);
#pragma omp parallel for schedule(static)
        for (size_t i = 0; i < L; i++)
            A.combined[i] = fadd(fadd(fmul(tnum[i], dz[i]), fmul(b0[i], d0[i])), fmul(b1[i], d1[i]));
    }

    // (3) composition polynomial: interpolate over the coset, split into 7 columns of n (App. A.6)
    std::vector<u128> comp_coeffs = A.combined;
    interpolate_poly_with_offset(comp_coeffs, o);
    if (poly_degree(comp_coeffs) >= C * n)
        throw ProveError(2, "constraint composition degree too high (trace does not satisfy the AIR)");
    A.comp_polys.assign(C, std::vector<u128>());
    A.comp_lde.assign(L * C, 0);
#pragma omp parallel for schedule(dynamic)
    for (size_t j = 0; j < C; j++) {
        A.comp_polys[j].assign(comp_coeffs.begin() + j * n, comp_coeffs.begin() + (j + 1) * n);
        auto ev = evaluate_poly_with_offset(A.comp_polys[j], o, B);
        for (size_t i = 0; i < L; i++) A.comp_lde[i * C + j] = ev[i];
    }
    MerkleTree comp_tree;
    {
        std::vector<Digest> leaves(L);
#pragma omp parallel for schedule(static)
        for (size_t i = 0; i < L; i++) leaves[i] = hash_elements(&A.comp_lde[i * C], C);
        comp_tree.build(std::move(leaves));
    }
    A.comp_root = comp_tree.root();
    commitments.insert(commitments.end(), A.comp_root.b, A.comp_root.b + 32);
    coin.reseed(A.comp_root);

    // (4) out-of-domain point and frame (App. A.7)
    A.z = coin.draw();
    const u128 zg = fmul(A.z, g);
    A.ood_cur.resize(W), A.ood_next.resize(W), A.ood_comp.resize(C);
    for (size_t c = 0; c < W; c++) {
        A.ood_cur[c] = eval_horner(A.trace_polys[c].data(), n, A.z);
        A.ood_next[c] = eval_horner(A.trace_polys[c].data(), n, zg);
    }
    std::vector<u128> ood_states;
    if (cp.ood_interleaved) {
        for (size_t c = 0; c < W; c++) ood_states.push_back(A.ood_cur[c]), ood_states.push_back(A.ood_next[c]);
    } else {
        ood_states = A.ood_cur;
        ood_states.insert(ood_states.end(), A.ood_next.begin(), A.ood_next.end());
    }
    coin.reseed(hash_elements(ood_states.data(), ood_states.size()));
    for (size_t j = 0; j < C; j++) A.ood_comp[j] = eval_horner(A.comp_polys[j].data(), n, A.z);
    coin.reseed(hash_elements(A.ood_comp.data(), C));

    // DEEP composition coefficients: 28 trace then 7 constraint draws (App. A.3 / A.8)
    A.deep_tc.resize(W), A.deep_cc.resize(C);
    for (auto& x : A.deep_tc) x = coin.draw();
    for (auto& x : A.deep_cc) x = coin.draw();

    // DEEP polynomial in coefficient space (App. A.8)
    std::vector<u128> deep(n, 0);
    {
        auto synth_div = [&](std::vector<u128>& p, u128 a) {  // p(x) / (x - a), exact division
            u128 carry = 0;
            for (size_t i = p.size(); i-- > 0;) {
                u128 t = fadd(p[i], fmul(carry, a));
                p[i] = carry;
                carry = t;
            }
        };
        std::vector<u128> t1(n, 0), t2(n, 0);
        for (size_t c = 0; c < W; c++)
            for (size_t m = 0; m < n; m++) t1[m] = fadd(t1[m], fmul(A.trace_polys[c][m], A.deep_tc[c]));
        t2 = t1;
        u128 s1 = 0, s2 = 0;
        for (size_t c = 0; c < W; c++) {
            s1 = fadd(s1, fmul(A.ood_cur[c], A.deep_tc[c]));
            s2 = fadd(s2, fmul(A.ood_next[c], A.deep_tc[c]));
        }
        t1[0] = fsub(t1[0], s1);
        t2[0] = fsub(t2[0], s2);
        synth_div(t1, A.z);
        synth_div(t2, zg);
        std::vector<u128> cq(n, 0);
        u128 s3 = 0;
        for (size_t j = 0; j < C; j++) {
            for (size_t m = 0; m < n; m++) cq[m] = fadd(cq[m], fmul(A.comp_polys[j][m], A.deep_cc[j]));
            s3 = fadd(s3, fmul(A.ood_comp[j], A.deep_cc[j]));
        }
        cq[0] = fsub(cq[0], s3);
        synth_div(cq, A.z);
        for (size_t m = 0; m < n; m++) deep[m] = fadd(fadd(t1[m], t2[m]), cq[m]);
    }
    A.deep_degree = poly_degree(deep);
    // winterfell: assert_eq!(trace_length - 2, deep_poly.degree())  (vm/src/processor/mod.rs:38-42 explains why)
    if (A.deep_degree != n - 2) throw ProveError(3, "DEEP composition degree != trace_length - 2");

    // (5) DEEP evaluations over the LDE domain
    A.deep_evals = evaluate_poly_with_offset(deep, o, B);

    // (6) FRI layers (App. A.9)
    std::vector<MerkleTree> fri_trees;
    std::vector<u128> evals = A.deep_evals;
    const size_t F = opt.fri_fold;
    const size_t nlayers = num_fri_layers(L, opt);
    const u128 inv8 = finv((u128)F), zeta_inv = finv(root_of_unity(ilog2(F)));
    for (size_t layer = 0; layer < nlayers; layer++) {
        size_t s = evals.size(), m = s / F;
        A.fri_layer_evals.push_back(evals);
        std::vector<Digest> leaves(m);
#pragma omp parallel for schedule(static)
        for (size_t i = 0; i < m; i++) {
            u128 row[8];
            for (size_t j = 0; j < F; j++) row[j] = evals[i + j * m];
            leaves[i] = hash_elements(row, F);
        }
        MerkleTree t;
        t.build(std::move(leaves));
        commitments.insert(commitments.end(), t.root().b, t.root().b + 32);
        coin.reseed(t.root());
        A.fri_roots.push_back(t.root());
        fri_trees.push_back(std::move(t));
        u128 alpha = coin.draw();
        A.fri_alphas.push_back(alpha);
        // degree-respecting projection: interpolate the 8 points (x*zeta^j, row[j]), evaluate at alpha
        std::vector<u128> nextv(m);
        u128 gs_inv = finv(root_of_unity(ilog2(s))), o_inv = finv(o);
        auto xinv_pows = power_table(gs_inv, m);
#pragma omp parallel for schedule(static)
        for (size_t i = 0; i < m; i++) {
            u128 xinv = fmul(o_inv, xinv_pows[i]);
            u128 acc = 0, apow = 1, xk = inv8;  // xk = (1/8) * x^-k
            for (size_t k = 0; k < F; k++) {
                u128 sum = 0, zjk = 1, zk = fexp(zeta_inv, k);
                for (size_t j = 0; j < F; j++) {
                    sum = fadd(sum, fmul(evals[i + j * m], zjk));
                    zjk = fmul(zjk, zk);
                }
                acc = fadd(acc, fmul(fmul(sum, xk), apow));
                apow = fmul(apow, alpha);
                xk = fmul(xk, xinv);
            }
            nextv[i] = acc;
        }
        evals.swap(nextv);
    }
    A.fri_layer_evals.push_back(evals);
    {
        std::vector<u128> rem = evals;
        interpolate_poly_with_offset(rem, o);
        size_t keep = rem.size() / B;
        for (size_t i = keep; i < rem.size(); i++)
            if (rem[i] != 0) throw ProveError(4, "FRI remainder degree too high");
        rem.resize(keep);
        A.remainder = rem;
        std::vector<u128> ordered = rem;
        if (!cp.remainder_low_to_high) std::reverse(ordered.begin(), ordered.end());
        Digest d = hash_elements(ordered.data(), ordered.size());
        commitments.insert(commitments.end(), d.b, d.b + 32);
        coin.reseed(d);
        A.fri_roots.push_back(d);
        A.remainder = ordered;
    }

    // (7) grinding + query positions
    uint64_t nonce = cp.first_nonce;
    while (coin.leading_zeros(nonce) < opt.grinding) nonce++;
    A.pow_nonce = nonce;
    auto positions = coin.draw_integers(opt.num_queries, L, nonce);
    std::sort(positions.begin(), positions.end());
    positions.erase(std::unique(positions.begin(), positions.end()), positions.end());
    A.positions = positions;

    // (8) assemble the proof (App. A.10)
    ByteWriter w;
    // Context: TraceInfo, modulus bytes, options
    w.u8((uint8_t)W), w.u8(0);                     // TraceInfo: main width, aux width
    if (cp.trace_info_aux_rands_byte) w.u8(0);     //            aux random elements
    w.u8((uint8_t)lgn), w.u16(0);                  //            log2(length), metadata length
    w.u8(16);
    w.elem(MOD);
    w.u8((uint8_t)opt.num_queries), w.u8((uint8_t)opt.blowup), w.u8((uint8_t)opt.grinding), w.u8((uint8_t)opt.field_ext);
    w.u8((uint8_t)opt.fri_fold), w.u8((uint8_t)opt.fri_rem_max_deg);
    w.u8((uint8_t)positions.size());
    w.u16((uint16_t)commitments.size());
    w.bytes(commitments.data(), commitments.size());
    auto write_queries = [&](const MerkleTree& t, const std::vector<size_t>& pos, const u128* table, size_t width) {
        std::vector<uint8_t> values;
        for (size_t p : pos) {
            const uint8_t* src = (const uint8_t*)&table[p * width];
            values.insert(values.end(), src, src + width * 16);
        }
        auto paths = serialize_nodes(prove_batch(t, pos));
        w.u32((uint32_t)values.size());
        w.bytes(values.data(), values.size());
        w.u32((uint32_t)paths.size());
        w.bytes(paths.data(), paths.size());
    };
    write_queries(trace_tree, positions, A.trace_lde.data(), W);
    write_queries(comp_tree, positions, A.comp_lde.data(), C);
    // OodFrame
    w.u16((uint16_t)(1 + ood_states.size() * 16));
    w.u8(2);
    for (u128 v : ood_states) w.elem(v);
    w.u16(1);
    w.u8(0);
    w.u16((uint16_t)(C * 16));
    for (u128 v : A.ood_comp) w.elem(v);
    // FriProof
    w.u8((uint8_t)nlayers);
    {
        std::vector<size_t> pos = positions;
        size_t domain = L;
        for (size_t layer = 0; layer < nlayers; layer++) {
            pos = fold_positions(pos, domain, F);
            size_t m = domain / F;
            const auto& ev = A.fri_layer_evals[layer];
            std::vector<uint8_t> values;
            for (size_t p : pos)
                for (size_t j = 0; j < F; j++) {
                    uint8_t t[16];
                    store_le(t, ev[p + j * m]);
                    values.insert(values.end(), t, t + 16);
                }
            auto paths = serialize_nodes(prove_batch(fri_trees[layer], pos));
            w.u32((uint32_t)values.size());
            w.bytes(values.data(), values.size());
            w.u32((uint32_t)paths.size());
            w.bytes(paths.data(), paths.size());
            domain = m;
        }
    }
    w.u16((uint16_t)(A.remainder.size() * 16));
    for (u128 v : A.remainder) w.elem(v);
    w.u8(1);  // num_partitions
    w.u64(nonce);
    w.u8(0);  // gkr_proof: None
    A.proof = std::move(w.b);
}

// ---------------------------------------------------------------------------------------------
// winter-verifier 0.9.0 `verify`, as called at vm/src/lib.rs:93-98 and
// examples/linear_regression/src/main.rs:85 with MinConjecturedSecurity(95).
struct ByteReader {
    const uint8_t* p;
    size_t len, pos = 0;
    bool ok = true;
    ByteReader(const uint8_t* p_, size_t l) : p(p_), len(l) {}
    bool need(size_t k) {
        if (pos + k > len) ok = false;
        return ok;
    }
    uint8_t u8() { return need(1) ? p[pos++] : 0; }
    uint16_t u16() {
        if (!need(2)) return 0;
        uint16_t v = p[pos] | (p[pos + 1] << 8);
        pos += 2;
        return v;
    }
    uint32_t u32() {
        if (!need(4)) return 0;
        uint32_t v = 0;
        for (int i = 0; i < 4; i++) v |= (uint32_t)p[pos + i] << (8 * i);
        pos += 4;
        return v;
    }
    uint64_t u64() {
        if (!need(8)) return 0;
        uint64_t v = 0;
        for (int i = 0; i < 8; i++) v |= (uint64_t)p[pos + i] << (8 * i);
        pos += 8;
        return v;
    }
    const uint8_t* bytes(size_t k) {
        if (!need(k)) return nullptr;
        const uint8_t* r = p + pos;
        pos += k;
        return r;
    }
    u128 elem() {  // winter-math: deserialization rejects values >= M
        const uint8_t* b = bytes(16);
        if (!b) return 0;
        u128 v = load_le(b);
        if (v >= MOD) ok = false;
        return v;
    }
};

enum VerifyResult {
    VERIFY_OK = 0,
    VERIFY_MALFORMED = 1,
    VERIFY_OPTIONS = 2,
    VERIFY_OOD_MISMATCH = 3,
    VERIFY_QUERY_COUNT = 4,
    VERIFY_TRACE_OPENING = 5,
    VERIFY_CONSTRAINT_OPENING = 6,
    VERIFY_FRI_OPENING = 7,
    VERIFY_FRI_FOLDING = 8,
    VERIFY_FRI_REMAINDER = 9,
    VERIFY_POW = 10,
    VERIFY_SECURITY = 11,
};

static inline int verify(const uint8_t* proof, size_t proof_len, const u128* pub18, const AirParams& ap,
                         unsigned min_conjectured_security, const Compat& cp) {
    ByteReader r(proof, proof_len);
    // Context
    unsigned W = r.u8(), aux = r.u8(), aux_rands = cp.trace_info_aux_rands_byte ? r.u8() : 0;
    unsigned lgn = r.u8(), meta = r.u16();
    unsigned modlen = r.u8();
    if (!r.ok || W != TRACE_WIDTH || aux != 0 || aux_rands != 0 || meta != 0 || modlen != 16 || lgn < 4 || lgn > 32) return VERIFY_MALFORMED;
    const uint8_t* modb = r.bytes(16);
    if (!modb || load_le(modb) != MOD) return VERIFY_MALFORMED;
    ProofOptions opt;
    opt.num_queries = r.u8(), opt.blowup = r.u8(), opt.grinding = r.u8(), opt.field_ext = r.u8();
    opt.fri_fold = r.u8(), opt.fri_rem_max_deg = r.u8();
    if (!r.ok || opt.field_ext != 1 || opt.fri_fold != 8 || opt.blowup != 8 || opt.num_queries == 0) return VERIFY_OPTIONS;
    const size_t n = (size_t)1 << lgn, B = opt.blowup, L = n * B, C = NUM_COMP_COLUMNS, F = opt.fri_fold;
    {   // conjectured security (App. A.2): min(min(128 - log2 L, queries*log2(blowup) [+ grinding]) - 1, 128);
        // the grinding bits only count once the queries alone give 80 bits (GRINDING_CONTRIBUTION_FLOOR)
        unsigned a = opt.num_queries * ilog2(B), b = 128, c = 128 - ilog2(L);
        if (a >= 80) a += opt.grinding;
        unsigned sec = std::min(std::min(a, c) - 1, b);
        if (sec < min_conjectured_security) return VERIFY_SECURITY;
    }
    unsigned num_unique = r.u8();
    size_t clen = r.u16();
    const size_t nlayers = num_fri_layers(L, opt);
    if (!r.ok || clen != 32 * (2 + nlayers + 1)) return VERIFY_MALFORMED;
    const uint8_t* cb = r.bytes(clen);
    if (!cb) return VERIFY_MALFORMED;
    auto commitment = [&](size_t i) {
        Digest d;
        memcpy(d.b, cb + 32 * i, 32);
        return d;
    };
    auto read_blob = [&](std::vector<uint8_t>& out) {
        size_t k = r.u32();
        const uint8_t* b = r.bytes(k);
        if (!b) return false;
        out.assign(b, b + k);
        return true;
    };
    std::vector<uint8_t> tq_values, tq_paths, cq_values, cq_paths;
    if (!read_blob(tq_values) || !read_blob(tq_paths) || !read_blob(cq_values) || !read_blob(cq_paths)) return VERIFY_MALFORMED;
    // OodFrame
    size_t ts_len = r.u16();
    if (!r.ok || ts_len != 1 + 2 * TRACE_WIDTH * 16) return VERIFY_MALFORMED;
    if (r.u8() != 2) return VERIFY_MALFORMED;
    std::vector<u128> ood_states(2 * TRACE_WIDTH);
    for (auto& v : ood_states) v = r.elem();
    if (r.u16() != 1 || r.u8() != 0) return VERIFY_MALFORMED;
    if (r.u16() != C * 16) return VERIFY_MALFORMED;
    std::vector<u128> ood_comp(C);
    for (auto& v : ood_comp) v = r.elem();
    // FriProof
    if (r.u8() != nlayers) return VERIFY_MALFORMED;
    std::vector<std::vector<uint8_t>> fl_values(nlayers), fl_paths(nlayers);
    for (size_t i = 0; i < nlayers; i++)
        if (!read_blob(fl_values[i]) || !read_blob(fl_paths[i])) return VERIFY_MALFORMED;
    size_t rem_len = r.u16();
    if (!r.ok || rem_len % 16) return VERIFY_MALFORMED;
    std::vector<u128> remainder(rem_len / 16);
    for (auto& v : remainder) v = r.elem();
    if (r.u8() != 1) return VERIFY_MALFORMED;
    uint64_t pow_nonce = r.u64();
    if (r.u8() != 0) return VERIFY_MALFORMED;
    if (!r.ok || r.pos != r.len) return VERIFY_MALFORMED;

    std::vector<u128> ood_cur(TRACE_WIDTH), ood_next(TRACE_WIDTH);
    for (size_t c = 0; c < TRACE_WIDTH; c++) {
        if (cp.ood_interleaved)
            ood_cur[c] = ood_states[2 * c], ood_next[c] = ood_states[2 * c + 1];
        else
            ood_cur[c] = ood_states[c], ood_next[c] = ood_states[TRACE_WIDTH + c];
    }

    // replay the transcript
    Coin coin;
    coin.init(coin_seed_elements(n, opt, pub18));
    coin.reseed(commitment(0));
    std::vector<u128> tcoef(NUM_TRANSITION), bcoef(NUM_ASSERTIONS);
    for (auto& x : tcoef) x = coin.draw();
    for (auto& x : bcoef) x = coin.draw();
    coin.reseed(commitment(1));
    const u128 z = coin.draw();
    const u128 o = GENERATOR, g = root_of_unity(lgn);

    // OOD consistency check
    {
        auto polys = periodic_polys();
        u128 zp = fexp(z, n / CYCLE_LENGTH), pv[NUM_PERIODIC];
        for (unsigned p = 0; p < NUM_PERIODIC; p++) pv[p] = eval_horner(polys[p].data(), CYCLE_LENGTH, zp);
        u128 ev[NUM_TRANSITION];
        evaluate_transition(ood_cur.data(), ood_next.data(), pv, ap, ev);
        u128 t = 0;
        for (unsigned j = 0; j < NUM_TRANSITION; j++) t = fadd(t, fmul(tcoef[j], ev[j]));
        u128 g_last = fexp(g, n - NUM_EXEMPTIONS), g_last2 = fexp(g, n - 1);
        u128 zt_num = fsub(fexp(z, n), 1), zt_den = fmul(fsub(z, g_last), fsub(z, g_last2));
        u128 result = fmul(fmul(t, zt_den), finv(zt_num));
        auto asserts = sorted_assertions(n, pub18);
        u128 s0 = 0, s1 = 0;
        for (unsigned k = 0; k < NUM_ASSERTIONS; k++) {
            u128 term = fmul(bcoef[k], fsub(ood_cur[asserts[k].column], asserts[k].value));
            if (asserts[k].step == 0)
                s0 = fadd(s0, term);
            else
                s1 = fadd(s1, term);
        }
        result = fadd(result, fmul(s0, finv(fsub(z, 1))));
        result = fadd(result, fmul(s1, finv(fsub(z, g_last))));
        coin.reseed(hash_elements(ood_states.data(), ood_states.size()));
        u128 rhs = 0;
        for (size_t j = 0; j < C; j++) rhs = fadd(rhs, fmul(fexp(z, (u128)j * n), ood_comp[j]));
        coin.reseed(hash_elements(ood_comp.data(), C));
        if (result != rhs) return VERIFY_OOD_MISMATCH;
    }
    std::vector<u128> deep_tc(TRACE_WIDTH), deep_cc(C);
    for (auto& x : deep_tc) x = coin.draw();
    for (auto& x : deep_cc) x = coin.draw();

    // FRI commitments -> alphas (one per commitment, incl. the remainder's)
    std::vector<u128> alphas;
    for (size_t i = 0; i < nlayers + 1; i++) {
        coin.reseed(commitment(2 + i));
        alphas.push_back(coin.draw());
    }
    if (coin.leading_zeros(pow_nonce) < opt.grinding) return VERIFY_POW;
    auto positions = coin.draw_integers(opt.num_queries, L, pow_nonce);
    std::sort(positions.begin(), positions.end());
    positions.erase(std::unique(positions.begin(), positions.end()), positions.end());
    if (positions.size() != num_unique) return VERIFY_QUERY_COUNT;
    const size_t Q = positions.size();

    // openings
    if (tq_values.size() != Q * TRACE_WIDTH * 16 || cq_values.size() != Q * C * 16) return VERIFY_MALFORMED;
    std::vector<u128> trows(Q * TRACE_WIDTH), crows(Q * C);
    {
        ByteReader tr(tq_values.data(), tq_values.size()), cr(cq_values.data(), cq_values.size());
        for (auto& v : trows) v = tr.elem();
        for (auto& v : crows) v = cr.elem();
        if (!tr.ok || !cr.ok) return VERIFY_MALFORMED;
    }
    {
        BatchProof bp;
        bp.depth = ilog2(L);
        if (!parse_nodes(tq_paths.data(), tq_paths.size(), bp.nodes)) return VERIFY_MALFORMED;
        for (size_t q = 0; q < Q; q++) bp.leaves.push_back(hash_elements(&trows[q * TRACE_WIDTH], TRACE_WIDTH));
        Digest root;
        if (!batch_root(bp, positions, root) || root != commitment(0)) return VERIFY_TRACE_OPENING;
    }
    {
        BatchProof bp;
        bp.depth = ilog2(L);
        if (!parse_nodes(cq_paths.data(), cq_paths.size(), bp.nodes)) return VERIFY_MALFORMED;
        for (size_t q = 0; q < Q; q++) bp.leaves.push_back(hash_elements(&crows[q * C], C));
        Digest root;
        if (!batch_root(bp, positions, root) || root != commitment(1)) return VERIFY_CONSTRAINT_OPENING;
    }

    // DEEP composition at the queried points (App. A.8 / A.11)
    const u128 wL = root_of_unity(ilog2(L)), zg = fmul(z, g);
    std::vector<u128> evaluations(Q);
    for (size_t q = 0; q < Q; q++) {
        u128 x = fmul(o, fexp(wL, positions[q]));
        u128 t1 = 0, t2 = 0;
        for (size_t c = 0; c < TRACE_WIDTH; c++) {
            u128 v = trows[q * TRACE_WIDTH + c];
            t1 = fadd(t1, fmul(fsub(v, ood_cur[c]), deep_tc[c]));
            t2 = fadd(t2, fmul(fsub(v, ood_next[c]), deep_tc[c]));
        }
        u128 d1 = fsub(x, z), d2 = fsub(x, zg);
        u128 tcomp = fmul(fadd(fmul(t1, d2), fmul(t2, d1)), finv(fmul(d1, d2)));
        u128 cnum = 0;
        for (size_t j = 0; j < C; j++) cnum = fadd(cnum, fmul(fsub(crows[q * C + j], ood_comp[j]), deep_cc[j]));
        evaluations[q] = fadd(tcomp, fmul(cnum, finv(d1)));
    }

    // FRI verification (App. A.9 / A.11)
    {
        std::vector<size_t> pos = positions;
        size_t domain = L;
        u128 dg = wL;
        size_t max_degree_plus_1 = n;
        const u128 zeta = root_of_unity(ilog2(F));
        for (size_t layer = 0; layer < nlayers; layer++) {
            auto folded = fold_positions(pos, domain, F);
            size_t m = domain / F;
            if (fl_values[layer].size() != folded.size() * F * 16) return VERIFY_MALFORMED;
            std::vector<u128> vals(folded.size() * F);
            ByteReader vr(fl_values[layer].data(), fl_values[layer].size());
            for (auto& v : vals) v = vr.elem();
            if (!vr.ok) return VERIFY_MALFORMED;
            BatchProof bp;
            bp.depth = ilog2(m);
            if (!parse_nodes(fl_paths[layer].data(), fl_paths[layer].size(), bp.nodes)) return VERIFY_MALFORMED;
            for (size_t k = 0; k < folded.size(); k++) bp.leaves.push_back(hash_elements(&vals[k * F], F));
            Digest root;
            if (!batch_root(bp, folded, root) || root != commitment(2 + layer)) return VERIFY_FRI_OPENING;
            // the values claimed by the previous layer must appear in this layer's rows
            for (size_t q = 0; q < pos.size(); q++) {
                size_t idx = std::find(folded.begin(), folded.end(), pos[q] % m) - folded.begin();
                if (vals[idx * F + pos[q] / m] != evaluations[q]) return VERIFY_FRI_FOLDING;
            }
            // fold each row: Lagrange-interpolate through (xe*zeta^j, v_j), evaluate at alpha
            std::vector<u128> nextv(folded.size());
            for (size_t k = 0; k < folded.size(); k++) {
                u128 xe = fmul(fexp(dg, folded[k]), o);
                u128 xs[8];
                for (size_t j = 0; j < F; j++) xs[j] = fmul(xe, fexp(zeta, j));
                u128 acc = 0;
                for (size_t j = 0; j < F; j++) {
                    u128 num = 1, den = 1;
                    for (size_t t = 0; t < F; t++) {
                        if (t == j) continue;
                        num = fmul(num, fsub(alphas[layer], xs[t]));
                        den = fmul(den, fsub(xs[j], xs[t]));
                    }
                    acc = fadd(acc, fmul(vals[k * F + j], fmul(num, finv(den))));
                }
                nextv[k] = acc;
            }
            if (max_degree_plus_1 % F != 0) return VERIFY_FRI_FOLDING;
            evaluations = nextv;
            pos = folded;
            dg = fexp(dg, F);
            max_degree_plus_1 /= F;
            domain = m;
        }
        if (remainder.size() > max_degree_plus_1) return VERIFY_FRI_REMAINDER;
        // remainder commitment check
        {
            Digest d = hash_elements(remainder.data(), remainder.size());
            if (d != commitment(2 + nlayers)) return VERIFY_FRI_REMAINDER;
        }
        std::vector<u128> rem = remainder;
        if (!cp.remainder_low_to_high) std::reverse(rem.begin(), rem.end());
        for (size_t q = 0; q < pos.size(); q++) {
            u128 x = fmul(o, fexp(dg, pos[q]));
            if (eval_horner(rem.data(), rem.size(), x) != evaluations[q]) return VERIFY_FRI_REMAINDER;
        }
    }
    return VERIFY_OK;
}

}  // namespace orc
