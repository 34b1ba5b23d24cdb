// Verifier for the proofs this backend (and the reference) produces: `winterfell::verify::<ProcessorAir,
// Blake3_256, DefaultRandomCoin>(proof, pub_inputs, &MinConjecturedSecurity(95))` as called at
// vm/src/lib.rs:91-98 and examples/linear_regression/src/main.rs:81-85 (next-row 8f-1).
//
// The transcript replay, Merkle openings, DEEP composition at the queried points and the FRI checks are host C++
// (kilobytes of data); the AIR's 20 transition constraints at the out-of-domain point are evaluated by the same
// CUDA code the prover uses (csrc/air/processor_air.cuh through evaluate_frames), so there is one AIR
// implementation on the product side.
#include "prover.h"
#include "../../include/ezkvm_prover.h"
#include "../../include/ezkvm_rescue_constants.h"
#include "host/air_host.h"
#include <algorithm>
#include <cstring>
#include <map>

namespace ezk {

namespace {

constexpr uint32_t kWidth = 28, kCompCols = 7, kTransitions = 20, kAssertions = 22;

struct Reader {
    const uint8_t* p;
    size_t len, pos = 0;
    bool ok = true;
    bool need(size_t k) {
        if (pos + k > len) ok = false;
        return ok;
    }
    uint64_t le(int bytes) {
        if (!need(bytes)) return 0;
        uint64_t v = 0;
        for (int i = 0; i < bytes; i++) v |= (uint64_t)p[pos + i] << (8 * i);
        pos += bytes;
        return v;
    }
    const uint8_t* bytes(size_t k) {
        if (!need(k)) return nullptr;
        const uint8_t* r = p + pos;
        pos += k;
        return r;
    }
    Fp elem() {  // winter-math rejects non-canonical encodings
        const uint8_t* b = bytes(16);
        if (!b) return Fp();
        Fp v = fp_load(b);
        if (v.v >= Fp::modulus()) ok = false;
        return v;
    }
    bool blob(std::vector<uint8_t>& out) {
        const size_t k = (size_t)le(4);
        const uint8_t* b = bytes(k);
        if (!b) return false;
        out.assign(b, b + k);
        return true;
    }
};

struct Reject {
    const char* why;
};

// Root of a batch Merkle proof: the serialized digests are matched to node-array indices with the same index walk
// the prover uses to write them (batch_proof_node_indices), then every known node is hashed up to node 1.
Hash32 batch_root(const std::vector<Hash32>& leaves, const std::vector<uint64_t>& positions, uint64_t num_leaves,
                  const std::vector<uint8_t>& ser) {
    const auto lists = batch_proof_node_indices(num_leaves, positions);
    Reader r{ser.data(), ser.size()};
    if (r.le(1) != lists.size()) throw Reject{"batch proof: wrong number of paths"};
    std::map<uint64_t, Hash32> known;
    for (size_t q = 0; q < positions.size(); q++) known[num_leaves + positions[q]] = leaves[q];
    for (const auto& list : lists) {
        if (r.le(1) != list.size()) throw Reject{"batch proof: wrong path length"};
        for (uint64_t idx : list) {
            const uint8_t* d = r.bytes(32);
            if (!d) throw Reject{"batch proof: truncated"};
            Hash32 h;
            memcpy(h.data(), d, 32);
            known[idx] = h;
        }
    }
    if (!r.ok || r.pos != r.len) throw Reject{"batch proof: trailing bytes"};
    // highest index first: children before parents
    while (!known.empty()) {
        auto it = std::prev(known.end());
        const uint64_t k = it->first;
        if (k == 1) return it->second;
        auto sib = known.find(k ^ 1);
        if (sib == known.end()) throw Reject{"batch proof: missing sibling"};
        const Hash32 parent = (k & 1) ? merge_digests(sib->second, it->second) : merge_digests(it->second, sib->second);
        known.erase(k);
        known.erase(k ^ 1);
        known[k >> 1] = parent;
    }
    throw Reject{"batch proof: empty"};
}

}  // namespace

void GpuProver::verify(const uint8_t* proof, size_t proof_len, const PublicInputs& pub, uint32_t min_security) {
    try {
        Reader r{proof, proof_len};
        // ---- Proof::context ----
        const WireCompat wc = wire_compat();
        const uint32_t W = (uint32_t)r.le(1), aux = (uint32_t)r.le(1), aux_rands = wc.trace_info_aux_rands_byte ? (uint32_t)r.le(1) : 0;
        const uint32_t log_n = (uint32_t)r.le(1), meta = (uint32_t)r.le(2);
        const uint32_t modlen = (uint32_t)r.le(1);
        const uint8_t* modb = r.bytes(16);
        if (!r.ok || W != kWidth || aux != 0 || aux_rands != 0 || meta != 0 || modlen != 16 || log_n < 6 || log_n > 32 ||
            fp_load(modb).v != Fp::modulus())
            throw Reject{"malformed proof context"};
        ProofOptions opt;
        opt.num_queries = (uint32_t)r.le(1), opt.blowup = (uint32_t)r.le(1), opt.grinding = (uint32_t)r.le(1);
        opt.field_ext = (uint32_t)r.le(1), opt.fri_fold = (uint32_t)r.le(1), opt.fri_rem_max_deg = (uint32_t)r.le(1);
        if (!r.ok || opt.field_ext != 1 || opt.fri_fold != 8 || opt.blowup != 8 || opt.num_queries == 0)
            throw Reject{"unsupported proof options"};
        const uint64_t n = 1ull << log_n, L = 8 * n;
        const uint32_t log_L = log_n + 3;
        {   // conjectured security: min(min(128 - log2 L, queries * log2(blowup) [+ grinding]) - 1, 128); grinding
            // bits count only once the queries alone reach 80 bits (winter-air's GRINDING_CONTRIBUTION_FLOOR)
            uint32_t query_bits = opt.num_queries * 3;
            if (query_bits >= 80) query_bits += opt.grinding;
            const uint32_t sec = std::min(std::min(query_bits, 128u - log_L) - 1, 128u);
            if (sec < min_security) throw Reject{"proof does not reach the required conjectured security level"};
        }
        const uint32_t num_unique = (uint32_t)r.le(1);
        const size_t clen = (size_t)r.le(2);
        const size_t nlayers = num_fri_layers(L, opt);
        const uint8_t* cb = r.bytes(clen);
        if (!r.ok || clen != 32 * (2 + nlayers + 1)) throw Reject{"malformed commitments"};
        auto commitment = [&](size_t i) {
            Hash32 h;
            memcpy(h.data(), cb + 32 * i, 32);
            return h;
        };
        std::vector<uint8_t> tq_values, tq_paths, cq_values, cq_paths;
        if (!r.blob(tq_values) || !r.blob(tq_paths) || !r.blob(cq_values) || !r.blob(cq_paths)) throw Reject{"malformed queries"};
        // ---- OodFrame ----
        if (r.le(2) != 1 + 2 * kWidth * 16 || r.le(1) != 2) throw Reject{"malformed out-of-domain frame"};
        std::vector<Fp> ood_states(2 * kWidth);
        for (auto& v : ood_states) v = r.elem();
        if (r.le(2) != 1 || r.le(1) != 0 || r.le(2) != kCompCols * 16) throw Reject{"malformed out-of-domain frame"};
        std::vector<Fp> ood_comp(kCompCols);
        for (auto& v : ood_comp) v = r.elem();
        // ---- FriProof ----
        if (r.le(1) != nlayers) throw Reject{"wrong number of FRI layers"};
        std::vector<std::vector<uint8_t>> fl_values(nlayers), fl_paths(nlayers);
        for (size_t i = 0; i < nlayers; i++)
            if (!r.blob(fl_values[i]) || !r.blob(fl_paths[i])) throw Reject{"malformed FRI layer"};
        const size_t rem_len = (size_t)r.le(2);
        if (!r.ok || rem_len % 16) throw Reject{"malformed FRI remainder"};
        std::vector<Fp> remainder(rem_len / 16);
        for (auto& v : remainder) v = r.elem();
        if (r.le(1) != 1) throw Reject{"malformed FRI proof"};
        const uint64_t pow_nonce = r.le(8);
        if (r.le(1) != 0 || !r.ok || r.pos != r.len) throw Reject{"malformed proof tail"};

        std::vector<Fp> ood_cur(kWidth), ood_next(kWidth);  // serialized interleaved [cur_c, next_c] (WireCompat)
        for (uint32_t c = 0; c < kWidth; c++) {
            if (wc.ood_interleaved)
                ood_cur[c] = ood_states[2 * c], ood_next[c] = ood_states[2 * c + 1];
            else
                ood_cur[c] = ood_states[c], ood_next[c] = ood_states[kWidth + c];
        }

        // ---- replay the transcript ----
        RandomCoin coin;
        coin.init(coin_seed(kWidth, n, opt, pub.elements));
        coin.reseed(commitment(0));
        std::vector<Fp> tcoef(kTransitions), bcoef(kAssertions);
        for (auto& x : tcoef) x = coin.draw();
        for (auto& x : bcoef) x = coin.draw();
        coin.reseed(commitment(1));
        const Fp z = coin.draw();
        const Fp o = Fp::from_u64(kDomainOffset), g = root_of_unity(log_n), zg = z * g;

        // ---- out-of-domain consistency: constraints(z) = sum_j z^(j n) H_j(z) ----
        {
            Fp pv[9];
            const auto polys = periodic_polys();
            const Fp zp = pow(z, n / 16);
            for (int p = 0; p < 9; p++) pv[p] = horner(polys[p], zp);
            Fp ev[kTransitions];
            stage_eval_frames(ood_cur.data(), ood_next.data(), pv, 1, pub.lwe_delta, ev);  // the AIR, on the GPU
            Fp t;
            for (uint32_t j = 0; j < kTransitions; j++) t = t + tcoef[j] * ev[j];
            const Fp g_last = pow(g, n - 2), g_last2 = pow(g, n - 1);
            Fp result = t * (z - g_last) * (z - g_last2) * inverse(pow(z, n) - Fp(1));
            // assertions sorted by (step, column): air/src/lib.rs:170-195
            const uint32_t cols0[12] = {0, 7, 8, 11, 12, 13, 14, 15, 16, 17, 18, 19};
            const uint32_t cols1[10] = {7, 8, 12, 13, 14, 15, 16, 17, 18, 19};
            Fp s0, s1;
            for (uint32_t k = 0; k < 12; k++) s0 = s0 + bcoef[k] * ood_cur[cols0[k]];
            for (uint32_t k = 0; k < 10; k++)
                s1 = s1 + bcoef[12 + k] * (ood_cur[cols1[k]] - (k < 2 ? pub.elements[k] : pub.elements[2 + (k - 2)]));
            result = result + s0 * inverse(z - Fp(1)) + s1 * inverse(z - g_last);
            Fp rhs, zn = pow(z, n), zj(1);
            for (uint32_t j = 0; j < kCompCols; j++) {
                rhs = rhs + zj * ood_comp[j];
                zj = zj * zn;
            }
            coin.reseed(hash_elements(ood_states.data(), ood_states.size()));
            coin.reseed(hash_elements(ood_comp.data(), ood_comp.size()));
            if (result != rhs) throw Reject{"out-of-domain constraint evaluations are inconsistent with the composition polynomial"};
        }
        std::vector<Fp> deep_tc(kWidth), deep_cc(kCompCols);
        for (auto& x : deep_tc) x = coin.draw();
        for (auto& x : deep_cc) x = coin.draw();
        std::vector<Fp> alphas;
        for (size_t i = 0; i < nlayers + 1; i++) {
            coin.reseed(commitment(2 + i));
            alphas.push_back(coin.draw());
        }
        if (coin.leading_zeros(pow_nonce) < opt.grinding) throw Reject{"proof-of-work nonce does not meet the grinding factor"};
        std::vector<uint64_t> positions = coin.draw_integers(opt.num_queries, L, pow_nonce);
        std::sort(positions.begin(), positions.end());
        positions.erase(std::unique(positions.begin(), positions.end()), positions.end());
        if (positions.size() != num_unique) throw Reject{"wrong number of unique query positions"};
        const size_t Q = positions.size();

        // ---- trace / constraint openings ----
        if (tq_values.size() != Q * kWidth * 16 || cq_values.size() != Q * kCompCols * 16) throw Reject{"malformed query values"};
        std::vector<Fp> trows(Q * kWidth), crows(Q * kCompCols);
        {
            Reader tr{tq_values.data(), tq_values.size()}, cr{cq_values.data(), cq_values.size()};
            for (auto& v : trows) v = tr.elem();
            for (auto& v : crows) v = cr.elem();
            if (!tr.ok || !cr.ok) throw Reject{"non-canonical query value"};
        }
        {
            std::vector<Hash32> leaves(Q);
            for (size_t q = 0; q < Q; q++) leaves[q] = hash_elements(&trows[q * kWidth], kWidth);
            if (batch_root(leaves, positions, L, tq_paths) != commitment(0)) throw Reject{"trace openings do not match the trace commitment"};
            for (size_t q = 0; q < Q; q++) leaves[q] = hash_elements(&crows[q * kCompCols], kCompCols);
            if (batch_root(leaves, positions, L, cq_paths) != commitment(1))
                throw Reject{"constraint openings do not match the constraint commitment"};
        }

        // ---- DEEP composition at the queried points ----
        const Fp wL = root_of_unity(log_L);
        std::vector<Fp> evaluations(Q);
        for (size_t q = 0; q < Q; q++) {
            const Fp x = o * pow(wL, positions[q]);
            Fp t1, t2, cnum;
            for (uint32_t c = 0; c < kWidth; c++) {
                const Fp v = trows[q * kWidth + c];
                t1 = t1 + (v - ood_cur[c]) * deep_tc[c];
                t2 = t2 + (v - ood_next[c]) * deep_tc[c];
            }
            for (uint32_t j = 0; j < kCompCols; j++) cnum = cnum + (crows[q * kCompCols + j] - ood_comp[j]) * deep_cc[j];
            const Fp d1 = x - z, d2 = x - zg;
            evaluations[q] = (t1 * d2 + t2 * d1) * inverse(d1 * d2) + cnum * inverse(d1);
        }

        // ---- FRI ----
        std::vector<uint64_t> pos = positions;
        uint64_t domain = L, max_degree_plus_1 = n;
        Fp dg = wL;
        const Fp zeta = root_of_unity(3);
        for (size_t layer = 0; layer < nlayers; layer++) {
            const std::vector<uint64_t> folded = fold_positions(pos, domain, 8);
            const uint64_t m = domain / 8;
            if (fl_values[layer].size() != folded.size() * 8 * 16) throw Reject{"malformed FRI layer values"};
            std::vector<Fp> vals(folded.size() * 8);
            Reader vr{fl_values[layer].data(), fl_values[layer].size()};
            for (auto& v : vals) v = vr.elem();
            if (!vr.ok) throw Reject{"non-canonical FRI value"};
            std::vector<Hash32> leaves(folded.size());
            for (size_t k = 0; k < folded.size(); k++) leaves[k] = hash_elements(&vals[k * 8], 8);
            if (batch_root(leaves, folded, m, fl_paths[layer]) != commitment(2 + layer))
                throw Reject{"FRI layer openings do not match the layer commitment"};
            for (size_t q = 0; q < pos.size(); q++) {
                const size_t idx = std::find(folded.begin(), folded.end(), pos[q] % m) - folded.begin();
                if (vals[idx * 8 + pos[q] / m] != evaluations[q]) throw Reject{"FRI layer values are inconsistent with the previous layer"};
            }
            // fold each opened row: interpolate through (x zeta^j, v_j), evaluate at alpha
            std::vector<Fp> next(folded.size());
            for (size_t k = 0; k < folded.size(); k++) {
                const Fp xe = pow(dg, folded[k]) * o;
                Fp xs[8], acc;
                for (int j = 0; j < 8; j++) xs[j] = xe * pow(zeta, j);
                for (int j = 0; j < 8; j++) {
                    Fp num(1), den(1);
                    for (int t = 0; t < 8; t++) {
                        if (t == j) continue;
                        num = num * (alphas[layer] - xs[t]);
                        den = den * (xs[j] - xs[t]);
                    }
                    acc = acc + vals[k * 8 + j] * num * inverse(den);
                }
                next[k] = acc;
            }
            if (max_degree_plus_1 % 8) throw Reject{"FRI degree bound is not divisible by the folding factor"};
            evaluations = next, pos = folded, dg = pow(dg, 8), max_degree_plus_1 /= 8, domain = m;
        }
        if (remainder.size() > max_degree_plus_1) throw Reject{"FRI remainder degree is too high"};
        if (hash_elements(remainder.data(), remainder.size()) != commitment(2 + nlayers))
            throw Reject{"FRI remainder does not match its commitment"};
        std::vector<Fp> rem_poly = remainder;  // the commitment covers the coefficients in wire order
        if (!wc.remainder_low_to_high) std::reverse(rem_poly.begin(), rem_poly.end());
        for (size_t q = 0; q < pos.size(); q++)
            if (horner(rem_poly, o * pow(dg, pos[q])) != evaluations[q]) throw Reject{"FRI remainder is inconsistent with the last layer"};
    } catch (const Reject& e) {
        throw ProveFailure{EZK_ERR_VERIFICATION, std::string("verification failed: ") + e.why};
    } catch (const std::runtime_error& e) {
        if (dynamic_cast<const CudaError*>(&e)) throw;
        throw ProveFailure{EZK_ERR_VERIFICATION, std::string("verification failed: ") + e.what()};
    }
}

}  // namespace ezk
