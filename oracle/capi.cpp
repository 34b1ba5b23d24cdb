// ORACLE (test infrastructure, not product): C entry points for ctypes.  Only tests/, the
// smoke check and bench.py's CPU-baseline / reference legs may load this library.
#include "stark.hpp"
#include <chrono>
#include <cstdio>

using namespace orc;

static thread_local std::string g_err;

extern "C" {

const char* orc_last_error() { return g_err.c_str(); }

int orc_num_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_set_num_threads(int t) {
#ifdef _OPENMP
    omp_set_num_threads(t);
#else
    (void)t;
#endif
}

// ---- field (elements are 16 little-endian bytes) ----
void orc_fadd(const u128* a, const u128* b, u128* out, size_t n) { for (size_t i = 0; i < n; i++) out[i] = fadd(a[i], b[i]); }
void orc_fsub(const u128* a, const u128* b, u128* out, size_t n) { for (size_t i = 0; i < n; i++) out[i] = fsub(a[i], b[i]); }
void orc_fmul(const u128* a, const u128* b, u128* out, size_t n) { for (size_t i = 0; i < n; i++) out[i] = fmul(a[i], b[i]); }
void orc_finv(const u128* a, u128* out, size_t n) { for (size_t i = 0; i < n; i++) out[i] = finv(a[i]); }
void orc_fexp(const u128* a, const u128* e, u128* out, size_t n) { for (size_t i = 0; i < n; i++) out[i] = fexp(a[i], e[i]); }
void orc_root_of_unity(unsigned log_n, u128* out) { *out = root_of_unity(log_n); }
void orc_batch_inverse(u128* v, size_t n) { batch_inverse(v, n); }

// ---- hash ----
void orc_blake3(const uint8_t* data, size_t len, uint8_t* out32) { b3::hash(data, len, out32); }
void orc_merge_with_int(const uint8_t* seed32, uint64_t v, uint8_t* out32) {
    Digest s;
    memcpy(s.b, seed32, 32);
    Digest d = merge_with_int(s, v);
    memcpy(out32, d.b, 32);
}

// ---- transforms ----
void orc_interpolate(u128* a, size_t n) {
    std::vector<u128> v(a, a + n);
    interpolate_poly(v);
    memcpy(a, v.data(), n * 16);
}
void orc_interpolate_with_offset(u128* a, size_t n) {
    std::vector<u128> v(a, a + n);
    interpolate_poly_with_offset(v, GENERATOR);
    memcpy(a, v.data(), n * 16);
}
void orc_forward_ntt(u128* a, size_t n) {
    std::vector<u128> v(a, a + n);
    forward_ntt(v);
    memcpy(a, v.data(), n * 16);
}
// coefficients (n) -> evaluations over 3*<w_{n*blowup}>
void orc_evaluate_with_offset(const u128* p, size_t n, size_t blowup, u128* out) {
    std::vector<u128> v(p, p + n);
    auto ev = evaluate_poly_with_offset(v, GENERATOR, blowup);
    memcpy(out, ev.data(), ev.size() * 16);
}
// column of n trace values -> L = n*blowup LDE values (iNTT then coset evaluation), and the coefficients
void orc_lde_column(const u128* col, size_t n, size_t blowup, u128* coeffs_out, u128* lde_out) {
    std::vector<u128> v(col, col + n);
    interpolate_poly(v);
    if (coeffs_out) memcpy(coeffs_out, v.data(), n * 16);
    auto ev = evaluate_poly_with_offset(v, GENERATOR, blowup);
    memcpy(lde_out, ev.data(), ev.size() * 16);
}
void orc_eval_horner(const u128* p, size_t n, const u128* x, u128* out) { *out = eval_horner(p, n, *x); }

// ---- Merkle ----
// rows: row-major num_rows x width elements; nodes_out: 2*num_rows digests (index 1 = root) or NULL
void orc_merkle_rows(const u128* rows, size_t num_rows, size_t width, uint8_t* root_out, uint8_t* nodes_out) {
    std::vector<Digest> leaves(num_rows);
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < num_rows; i++) leaves[i] = hash_elements(rows + i * width, width);
    MerkleTree t;
    t.build(std::move(leaves));
    memcpy(root_out, t.root().b, 32);
    if (nodes_out) memcpy(nodes_out, t.nodes.data(), t.nodes.size() * 32);
}
// returns serialized nodes length; leaves (hashed rows) are given as the tree's node array
size_t orc_prove_batch(const uint8_t* nodes, size_t num_leaves, const uint64_t* idx, size_t nidx, uint8_t* out, size_t cap) {
    MerkleTree t;
    t.num_leaves = num_leaves;
    t.depth = ilog2(num_leaves);
    t.nodes.resize(2 * num_leaves);
    memcpy(t.nodes.data(), nodes, 2 * num_leaves * 32);
    std::vector<size_t> ix(idx, idx + nidx);
    auto ser = serialize_nodes(prove_batch(t, ix));
    if (ser.size() <= cap) memcpy(out, ser.data(), ser.size());
    return ser.size();
}
// verifier side: recompute the root from leaves (input order) + serialized nodes; 0 on success
int orc_batch_root(const uint8_t* leaves, const uint64_t* idx, size_t nidx, unsigned depth, const uint8_t* ser,
                   size_t ser_len, uint8_t* root_out) {
    BatchProof bp;
    bp.depth = depth;
    if (!parse_nodes(ser, ser_len, bp.nodes)) return 1;
    bp.leaves.resize(nidx);
    memcpy(bp.leaves.data(), leaves, nidx * 32);
    std::vector<size_t> ix(idx, idx + nidx);
    Digest root;
    if (!batch_root(bp, ix, root)) return 2;
    memcpy(root_out, root.b, 32);
    return 0;
}

// ---- AIR ----
void orc_evaluate_transition(const u128* cur, const u128* nxt, const u128* periodic9, uint32_t lwe_k, uint32_t delta, u128* out20) {
    AirParams ap{lwe_k, delta};
    evaluate_transition(cur, nxt, periodic9, ap, out20);
}
// 9 x 16 values, column-major (column p at out[p*16 .. p*16+16])
void orc_periodic_columns(u128* out) {
    auto c = periodic_columns();
    for (unsigned p = 0; p < NUM_PERIODIC; p++) memcpy(out + p * CYCLE_LENGTH, c[p].data(), CYCLE_LENGTH * 16);
}
void orc_rescue_constants(u128* mds16, u128* inv_mds16, u128* ark128) {
    for (int i = 0; i < 16; i++) mds16[i] = c128(MDS_RAW[i]), inv_mds16[i] = c128(INV_MDS_RAW[i]);
    for (int i = 0; i < 128; i++) ark128[i] = c128(ARK_RAW[i]);
}
// winterfell's debug-build `trace.validate(&air)`: transitions on rows 0..n-3 and the 22 assertions.
// Returns -1 when valid, else row*64 + constraint index (assertion failures: 32 + k).
long orc_validate_trace(const u128* const* cols, size_t n, const u128* pub18, uint32_t lwe_k, uint32_t delta) {
    AirParams ap{lwe_k, delta};
    auto pc = periodic_columns();
    for (auto& a : sorted_assertions(n, pub18))
        if (cols[a.column][a.step] != a.value) return (long)(a.step * 64 + 32 + a.column);
    for (size_t i = 0; i + NUM_EXEMPTIONS < n; i++) {
        u128 cur[TRACE_WIDTH], nxt[TRACE_WIDTH], per[NUM_PERIODIC], ev[NUM_TRANSITION];
        for (unsigned c = 0; c < TRACE_WIDTH; c++) cur[c] = cols[c][i], nxt[c] = cols[c][i + 1];
        for (unsigned p = 0; p < NUM_PERIODIC; p++) per[p] = pc[p][i % CYCLE_LENGTH];
        evaluate_transition(cur, nxt, per, ap, ev);
        for (unsigned j = 0; j < NUM_TRANSITION; j++)
            if (ev[j] != 0) return (long)(i * 64 + j);
    }
    return -1;
}

// ---- prover / verifier ----
struct OrcOptions {
    uint32_t num_queries, blowup, grinding, field_ext, fri_fold, fri_rem_max_deg;
    uint32_t lwe_k, delta;
    uint32_t compat_ood_interleaved, compat_remainder_low_to_high, compat_trace_info_aux_rands_byte;
    uint64_t compat_first_nonce;
};

static void split(const OrcOptions* o, ProofOptions& po, AirParams& ap, Compat& cp) {
    po.num_queries = o->num_queries, po.blowup = o->blowup, po.grinding = o->grinding, po.field_ext = o->field_ext;
    po.fri_fold = o->fri_fold, po.fri_rem_max_deg = o->fri_rem_max_deg;
    ap.lwe_k = o->lwe_k, ap.delta = o->delta;
    cp.ood_interleaved = o->compat_ood_interleaved != 0;
    cp.remainder_low_to_high = o->compat_remainder_low_to_high != 0;
    cp.trace_info_aux_rands_byte = o->compat_trace_info_aux_rands_byte != 0;
    cp.first_nonce = o->compat_first_nonce;
}

// Returns an Artifacts handle (NULL on error; see orc_last_error). seconds_out: wall time of prove().
void* orc_prove(const u128* const* cols, size_t n, const u128* pub18, const OrcOptions* o, int* err_code, double* seconds_out) {
    ProofOptions po;
    AirParams ap;
    Compat cp;
    split(o, po, ap, cp);
    Artifacts* A = new Artifacts();
    if (err_code) *err_code = 0;
    try {
        auto t0 = std::chrono::steady_clock::now();
        prove(cols, n, pub18, ap, po, cp, *A);
        auto t1 = std::chrono::steady_clock::now();
        if (seconds_out) *seconds_out = std::chrono::duration<double>(t1 - t0).count();
    } catch (const ProveError& e) {
        g_err = e.what();
        if (err_code) *err_code = e.code;
        delete A;
        return nullptr;
    } catch (const std::exception& e) {
        g_err = e.what();
        if (err_code) *err_code = -1;
        delete A;
        return nullptr;
    }
    return A;
}
void orc_free_artifacts(void* h) { delete (Artifacts*)h; }

// field ids for orc_art_bytes / orc_art_copy
enum {
    ART_PROOF = 0, ART_TRACE_ROOT, ART_COMP_ROOT, ART_TCOEF, ART_BCOEF, ART_COMBINED, ART_Z, ART_OOD_CUR, ART_OOD_NEXT,
    ART_OOD_COMP, ART_DEEP_TC, ART_DEEP_CC, ART_DEEP_EVALS, ART_FRI_ROOTS, ART_FRI_ALPHAS, ART_REMAINDER, ART_POSITIONS,
    ART_TRACE_LDE, ART_COMP_LDE, ART_TRACE_POLYS, ART_COMP_POLYS, ART_POW_NONCE, ART_FRI_LAYER_EVALS
};

static std::vector<uint8_t> art_bytes(const Artifacts& A, int id, int sub) {
    auto elems = [](const std::vector<u128>& v) { return std::vector<uint8_t>((const uint8_t*)v.data(), (const uint8_t*)(v.data() + v.size())); };
    switch (id) {
        case ART_PROOF: return A.proof;
        case ART_TRACE_ROOT: return std::vector<uint8_t>(A.trace_root.b, A.trace_root.b + 32);
        case ART_COMP_ROOT: return std::vector<uint8_t>(A.comp_root.b, A.comp_root.b + 32);
        case ART_TCOEF: return elems(A.tcoef);
        case ART_BCOEF: return elems(A.bcoef);
        case ART_COMBINED: return elems(A.combined);
        case ART_Z: return elems(std::vector<u128>{A.z});
        case ART_OOD_CUR: return elems(A.ood_cur);
        case ART_OOD_NEXT: return elems(A.ood_next);
        case ART_OOD_COMP: return elems(A.ood_comp);
        case ART_DEEP_TC: return elems(A.deep_tc);
        case ART_DEEP_CC: return elems(A.deep_cc);
        case ART_DEEP_EVALS: return elems(A.deep_evals);
        case ART_FRI_ROOTS: {
            std::vector<uint8_t> out;
            for (auto& d : A.fri_roots) out.insert(out.end(), d.b, d.b + 32);
            return out;
        }
        case ART_FRI_ALPHAS: return elems(A.fri_alphas);
        case ART_REMAINDER: return elems(A.remainder);
        case ART_POSITIONS: {
            std::vector<uint64_t> p(A.positions.begin(), A.positions.end());
            return std::vector<uint8_t>((const uint8_t*)p.data(), (const uint8_t*)(p.data() + p.size()));
        }
        case ART_TRACE_LDE: return elems(A.trace_lde);
        case ART_COMP_LDE: return elems(A.comp_lde);
        case ART_TRACE_POLYS: return elems(A.trace_polys.at(sub));
        case ART_COMP_POLYS: return elems(A.comp_polys.at(sub));
        case ART_POW_NONCE: {
            uint64_t v = A.pow_nonce;
            return std::vector<uint8_t>((const uint8_t*)&v, (const uint8_t*)&v + 8);
        }
        case ART_FRI_LAYER_EVALS: return elems(A.fri_layer_evals.at(sub));
    }
    return {};
}
size_t orc_art_size(void* h, int id, int sub) { return art_bytes(*(Artifacts*)h, id, sub).size(); }
void orc_art_copy(void* h, int id, int sub, uint8_t* dst) {
    auto b = art_bytes(*(Artifacts*)h, id, sub);
    memcpy(dst, b.data(), b.size());
}

int orc_verify(const uint8_t* proof, size_t len, const u128* pub18, const OrcOptions* o, unsigned min_security) {
    ProofOptions po;
    AirParams ap;
    Compat cp;
    split(o, po, ap, cp);
    try {
        return verify(proof, len, pub18, ap, min_security, cp);
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}

}  // extern "C"
