// Host half of the prover: Fiat-Shamir transcript, query-position logic, batch Merkle proof index walk
// and proof serialization (kilobytes of data, but byte-exactness lives here).
// Mirrors winter-crypto's `Blake3_256` + `DefaultRandomCoin` (prover/src/lib.rs:13,44-45),
// winter-prover's ProverChannel and winter-air's `Proof` layout.
#pragma once
#include "../field/f128_host.h"
#include <array>
#include <cstdint>
#include <vector>

namespace ezk {

typedef std::array<uint8_t, 32> Hash32;

// BLAKE3 (default mode, 32-byte output) of an arbitrary-length message
Hash32 blake3(const uint8_t* data, size_t len);
Hash32 hash_elements(const Fp* e, size_t n);
Hash32 merge_digests(const Hash32& a, const Hash32& b);
Hash32 merge_with_int(const Hash32& seed, uint64_t v);

class RandomCoin {
public:
    void init(const std::vector<Fp>& seed_elements);
    void reseed(const Hash32& d);
    Fp draw();
    unsigned leading_zeros(uint64_t nonce) const;
    std::vector<uint64_t> draw_integers(size_t count, uint64_t domain_size, uint64_t nonce);

private:
    Hash32 next();
    Hash32 seed_{};
    uint64_t counter_ = 0;
};

struct ProofOptions {  // ProofOptions::new(32, 8, 0, FieldExtension::None, 8, 127) at vm/src/lib.rs:20
    uint32_t num_queries = 32, blowup = 8, grinding = 0, field_ext = 1, fri_fold = 8, fri_rem_max_deg = 127;
};

// Byte-level details of winter-* 0.9.0 that could not be checked against the crate itself: the
// proof writer (prover.cu) and the verifier (verifier.cu) read every one of them from this single process-wide
// struct, which mirrors the test oracle's `Compat` struct field by field, so the first real Winterfell proof can settle
// them in one place (ezk_set_wire_compat).  Defaults = the current reading of the 0.9.0 sources.
struct WireCompat {
    bool ood_interleaved = true;            // A.7: trace states hashed / serialized as [cur_0, next_0, cur_1, next_1, ...]
    bool remainder_low_to_high = true;      // A.9: FRI remainder coefficients, constant term first
    bool trace_info_aux_rands_byte = true;  // A.10: TraceInfo carries the aux segment's random-element count (u8)
    uint64_t first_nonce = 1;               // A.3 (7): the grinding search starts at 1
};
WireCompat& wire_compat();

size_t num_fri_layers(uint64_t lde_size, const ProofOptions& o);

// Context::to_elements() ++ PublicInputs::to_elements() (air/src/lib.rs:38-47)
std::vector<Fp> coin_seed(uint32_t trace_width, uint64_t trace_len, const ProofOptions& o, const Fp pub_inputs[18]);

std::vector<uint64_t> fold_positions(const std::vector<uint64_t>& positions, uint64_t domain_size, uint64_t folding);

// Index walk of MerkleTree::prove_batch: for each normalized leaf pair, the node-array indices (leaves live at
// num_leaves + i) whose digests go into the proof, in serialization order.
std::vector<std::vector<uint64_t>> batch_proof_node_indices(uint64_t num_leaves, const std::vector<uint64_t>& leaf_indexes);

class ProofWriter {
public:
    void u8(uint8_t v) { buf_.push_back(v); }
    void u16(uint16_t v);
    void u32(uint32_t v);
    void u64(uint64_t v);
    void bytes(const void* p, size_t n);
    void element(Fp v);
    std::vector<uint8_t>& data() { return buf_; }

private:
    std::vector<uint8_t> buf_;
};

}  // namespace ezk
