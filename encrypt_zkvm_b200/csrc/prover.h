// GPU prover: everything `winterfell::Prover::prove` does after trace generation, on one B200.
// Drop-in for `ExecutionProver::prove` (prover/src/lib.rs:40-77, called at vm/src/lib.rs:26).
#pragma once
#include "air/constraints.cuh"
#include "dist/comm.h"
#include "host/copy_pool.h"
#include "host/transcript.h"
#include "ntt/ntt.cuh"
#include <memory>
#include <string>
#include <vector>

namespace ezk {

struct ProveFailure {
    int code;  // ezk_status
    std::string message;
};

struct PublicInputs {
    Fp elements[18];  // program_hash[2] ++ stack_outputs[16]   (air/src/lib.rs:38-47)
    uint32_t lwe_k = 4, lwe_delta = 16;
};

// Executed operation list of the program (one code byte per operation, vm/src/processor/opcodes.rs:30-43) and the
// random last row: enough to build 8 of the 28 trace columns on the device (csrc/trace/expand.cuh).
struct OpList {
    const uint8_t* codes = nullptr;
    uint64_t count = 0;
    const uint8_t* last_row = nullptr;  // 28 x 16 bytes
};

class GpuProver {
public:
    explicit GpuProver(int device);
    ~GpuProver();
    GpuProver(const GpuProver&) = delete;

    // host columns (28 pointers) or device-resident trace; exactly one of them non-null
    // ops (host columns only): the columns of kOpColumns are generated on the device and their pointers are ignored
    std::vector<uint8_t> prove(const uint8_t* const* host_columns, const uint4* device_trace, uint64_t n,
                               const PublicInputs& pub, const ProofOptions& opt, const OpList* ops = nullptr);

    // Multi-GPU single proof: after join(), prove() must be called by every rank of the group with the
    // same trace, public inputs and options; each rank extends / evaluates / hashes the LDE cosets it owns, the
    // per-row products are all-gathered over NCCL, and every rank returns the same proof bytes.
    void join(int rank, int world, const uint8_t unique_id[128]);
    // the same with an in-process group (several provers of one process, one host thread each): tests on one GPU
    void join_local(int rank, LocalGroup* group);
    int world() const { return comm_.world(); }

    // winterfell::verify for this AIR (vm/src/lib.rs:91-98); throws ProveFailure{EZK_ERR_VERIFICATION} on rejection
    void verify(const uint8_t* proof, size_t proof_len, const PublicInputs& pub, uint32_t min_conjectured_security);

    const float* stage_ms() const { return stage_ms_; }
    void timer_start();
    float timer_stop();
    std::vector<uint8_t> artifact(int which);

    // stage-level helpers (host buffers)
    void stage_lde(const void* columns, uint32_t width, uint64_t n, void* lde_out);
    void stage_ntt(const void* columns, uint32_t width, uint64_t n, bool inverse, void* out);
    void stage_merkle(const void* table, uint32_t width, uint64_t rows, void* nodes_out);
    void stage_fri_fold(const void* evals, uint64_t s, Fp alpha, void* next_out);
    void stage_eval_frames(const void* cur, const void* nxt, const void* periodic, uint32_t nframes, uint32_t delta,
                           void* out20);
    void stage_eval_frames_sum(const void* cur, const void* nxt, const void* periodic, uint32_t nframes, uint32_t delta,
                               const Fp tcoef[20], void* out1);
    void bench_lde_merkle(uint32_t width, uint64_t n, int iters, float* lde_ms, float* merkle_ms);
    void bench_fri(uint64_t n, int iters, float* fri_ms);

private:
    std::vector<uint8_t> prove_impl(const uint8_t* const* host_columns, const uint4* device_trace, uint64_t n,
                                    const PublicInputs& pub, const ProofOptions& opt, const OpList* ops);
    struct Arena {
        uint4* base = nullptr;
        size_t capacity = 0, used = 0;  // in 16-byte units
    };
    uint4* alloc(size_t elems);
    void reserve(size_t elems);
    void reset_arena() { arena_.used = 0; }
    void sync();

    int device_;
    cudaStream_t stream_ = nullptr;
    cudaStream_t copy_stream_ = nullptr;  // host -> device trace upload, overlapped with the first transforms
    cudaEvent_t copy_ev_[16];
    cudaEvent_t col_ev_[32];              // host trace: upload of column c done / transforms of launch group g queued behind it
    cudaEvent_t grp_ev_[32];
    cudaStream_t aux_stream_ = nullptr;   // small independent kernels overlapped with latency-bound phases
    cudaEvent_t aux_ev_[4];
    cudaEvent_t share_ev_[64];            // multi-GPU: interpolation of a column round done / its all-gather done
    NttTables tables_;
    Comm comm_;
    Arena arena_;
    uint64_t arena_generation_ = 0;   // changes with every (re)allocation: peers re-open their mapping of it
    uint32_t arena_allocations_ = 0;
    uint8_t* pinned_ = nullptr;  // small host staging buffer
    size_t pinned_bytes_ = 0;
    // Staged trace upload for pageable caller memory (default; EZK_STAGED_UPLOAD=0 disables): a ring of page-locked slots
    // filled by a few host threads, drained by the copy stream.  Created on first use.
    static constexpr int kStageSlots = 3;
    static constexpr size_t kStageSlotBytes = 16u << 20;
    size_t stage_slot_bytes_ = kStageSlotBytes;
    uint8_t* stage_[kStageSlots] = {nullptr, nullptr, nullptr};
    cudaEvent_t stage_ev_[kStageSlots] = {nullptr, nullptr, nullptr};
    std::unique_ptr<CopyPool> copy_pool_;
    bool use_staged_upload(const uint8_t* const* host_columns);
    void staged_copy_column(uint4* d_dst, const uint8_t* src, size_t bytes, uint64_t& chunk);
    ConstraintParams* d_params_ = nullptr;
    uint32_t* d_flag_ = nullptr;
    cudaEvent_t ev_[16];
    cudaEvent_t timer_ev_[2];
    float stage_ms_[8] = {0};
    std::vector<Fp> ptable_;        // periodic columns over the LDE domain, cached per trace length
    uint32_t ptable_log_n_ = 0;
    // 1 / ((x - 1)(x - g^(n-2))) over this rank's LDE rows: the boundary-constraint divisors depend on the trace
    // length (and the row shard) only, so they are kept across proofs (L * 16 bytes; recomputed per proof in the
    // arena when the allocation fails)
    uint4* d_bden_ = nullptr;
    uint64_t bden_key_ = 0;

    // state of the last proof (device pointers into the arena + host copies), for ezk_prover_artifact
    struct Last {
        uint64_t n = 0, L = 0;
        uint4 *tlde = nullptr, *clde = nullptr, *tcoef = nullptr, *combined_copy = nullptr, *deep = nullptr;
        Hash32 trace_root{}, comp_root{};
        std::vector<Fp> ood_trace, ood_comp, remainder;
        std::vector<Hash32> fri_roots;
        std::vector<uint64_t> positions;
        bool keep_combined = false;
    } last_;
};

}  // namespace ezk
