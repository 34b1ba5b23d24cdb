// GPU prover orchestration: one stream, device workspace, host transcript.  See prover.h.
// Stage order follows winter-prover 0.9.0 `Prover::prove`.
#include "prover.h"
#include <cstdlib>
#include <unistd.h>
#include <thread>
#include "../../include/ezkvm_prover.h"
#include "../../include/ezkvm_rescue_constants.h"
#include "common.h"
#include "compose/compose.cuh"
#include "dist/shard_layout.h"
#include "fri/fri.cuh"
#include "host/launch_groups.h"
#include "host/air_host.h"
#include "merkle/merkle.cuh"
#include "trace/expand.cuh"
#include <algorithm>
#include <atomic>
#include <cstring>
#include <mutex>

namespace ezk {

namespace {
std::atomic<uint64_t> g_launches{0};

constexpr uint32_t kWidth = 28, kCompCols = 7, kTransitions = 20, kAssertions = 22, kCycle = 16, kPeriodic = 9;

inline void put(uint64_t dst[2], Fp v) {
    dst[0] = (uint64_t)v.v;
    dst[1] = (uint64_t)(v.v >> 64);
}
// precomputed form of a constant multiplier: c * 2^(32 i), i = 0..3 (fe_pre in f128.cuh)
inline void put_pre(uint64_t dst[4][2], Fp v) {
    const Fp two32 = Fp((u128)1 << 32);
    for (int i = 0; i < 4; i++) {
        put(dst[i], v);
        v = v * two32;
    }
}

}  // namespace

void count_launch(uint64_t k) { g_launches.fetch_add(k, std::memory_order_relaxed); }
uint64_t launch_count() { return g_launches.load(std::memory_order_relaxed); }

namespace {
struct ProfileSlot {
    cudaEvent_t start, stop;
    int id;
    uint64_t bytes;
};
struct Profiler {
    bool enabled = false;
    std::vector<ProfileSlot> slots;
    size_t used = 0;
    std::mutex mu;
} g_prof;
const char* const kKernelNames[K_COUNT] = {"ntt_strided_pass", "ntt_final_pass", "hash_rows", "merkle_level", "merkle_top",
                                            "gather", "pair_inverse", "constraints", "frames", "eval_polys", "deep_combine",
                                            "deep_pointwise", "all_zero", "fri_fold", "fri_remainder"};
}  // namespace

const char* kernel_name(int id) { return id >= 0 && id < K_COUNT ? kKernelNames[id] : "?"; }
void profile_enable(bool on) { g_prof.enabled = on; }
void profile_reset() {
    std::lock_guard<std::mutex> lock(g_prof.mu);
    g_prof.used = 0;
}
void profile_read(int id, uint64_t* launches, double* ms, uint64_t* algo_bytes) {
    std::lock_guard<std::mutex> lock(g_prof.mu);
    cudaDeviceSynchronize();
    uint64_t n = 0, b = 0;
    double t = 0;
    for (size_t i = 0; i < g_prof.used; i++) {
        if (g_prof.slots[i].id != id) continue;
        float e = 0;
        if (cudaEventElapsedTime(&e, g_prof.slots[i].start, g_prof.slots[i].stop) == cudaSuccess) t += e;
        n++, b += g_prof.slots[i].bytes;
    }
    if (launches) *launches = n;
    if (ms) *ms = t;
    if (algo_bytes) *algo_bytes = b;
}

LaunchScope::LaunchScope(cudaStream_t s, KernelId id, uint64_t algo_bytes) : stream(s), slot(-1) {
    count_launch();
    if (!g_prof.enabled) return;
    std::lock_guard<std::mutex> lock(g_prof.mu);
    if (g_prof.used == g_prof.slots.size()) {
        ProfileSlot ps{};
        if (cudaEventCreate(&ps.start) != cudaSuccess || cudaEventCreate(&ps.stop) != cudaSuccess) return;
        g_prof.slots.push_back(ps);
    }
    slot = (int)g_prof.used++;
    g_prof.slots[slot].id = id;
    g_prof.slots[slot].bytes = algo_bytes;
    cudaEventRecord(g_prof.slots[slot].start, stream);
}
LaunchScope::~LaunchScope() {
    if (slot >= 0) cudaEventRecord(g_prof.slots[slot].stop, stream);
}

GpuProver::GpuProver(int device) : device_(device) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
        throw ProveFailure{EZK_ERR_NO_DEVICE, "no CUDA device available (this backend has no CPU fallback)"};
    if (device < 0 || device >= count) throw ProveFailure{EZK_ERR_INVALID_ARGUMENT, "invalid device ordinal"};
    EZK_CUDA(cudaSetDevice(device));
    EZK_CUDA(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
    ntt_tables_init(tables_);
    pinned_bytes_ = 8 << 20;  // query openings of every rank of a group (8 x 512 KiB), parameters, small read-backs
    EZK_CUDA(cudaMallocHost(&pinned_, pinned_bytes_));
    EZK_CUDA(cudaMalloc(&d_params_, sizeof(ConstraintParams)));
    EZK_CUDA(cudaMalloc(&d_flag_, sizeof(uint32_t)));
    for (auto& e : ev_) EZK_CUDA(cudaEventCreate(&e));
    for (auto& e : timer_ev_) EZK_CUDA(cudaEventCreate(&e));
    EZK_CUDA(cudaStreamCreateWithFlags(&copy_stream_, cudaStreamNonBlocking));
    {
        // the auxiliary stream carries the NCCL exchanges of the sharded proof and small kernels that must slip in
        // beside long transforms: highest priority, so its CTAs are scheduled as soon as the compute stream retires some
        int lo = 0, hi = 0;
        EZK_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        EZK_CUDA(cudaStreamCreateWithPriority(&aux_stream_, cudaStreamNonBlocking, hi));
    }
    for (auto& e : aux_ev_) EZK_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto& e : copy_ev_) EZK_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto& e : col_ev_) EZK_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto& e : grp_ev_) EZK_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto& e : share_ev_) EZK_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
}

// The overlap of the trace upload with the first transforms needs page-locked source memory; a drop-in caller
// (TraceTable's Vec<BaseElement> columns, vm/src/lib.rs:18) has pageable memory, for which cudaMemcpyAsync stages
// through one driver thread (41 ms for the 448 MiB of a 2^20-row trace, measured: longer than the proof itself).  Such
// columns therefore go through the prover's own page-locked ring, filled by a few host threads (copy_pool.h); memory
// the caller has page-locked (cudaHostRegister / cudaMallocHost) or managed keeps the plain asynchronous copy.
// EZK_STAGED_UPLOAD=0 switches the ring off (measurement knob).
bool GpuProver::use_staged_upload(const uint8_t* const* host_columns) {
    const char* v = getenv("EZK_STAGED_UPLOAD");  // read per proof: tests switch it inside one process
    if (v && v[0] == '0') return false;
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, host_columns[7]) != cudaSuccess) {  // column 7 always comes from the caller
        cudaGetLastError();  // older drivers report unregistered memory as an error: that is the pageable case
    } else if (attr.type != cudaMemoryTypeUnregistered) {
        return false;  // page-locked or managed: the plain asynchronous copy already overlaps
    }
    if (!copy_pool_) {
        const unsigned hw = std::thread::hardware_concurrency();
        unsigned threads = std::max(1u, std::min(8u, hw / 2));  // pieces are claimed dynamically: late threads cost nothing
        if (const char* t = getenv("EZK_STAGE_THREADS")) {  // measurement knob (tools/pageable_e2e.py)
            const long k = atol(t);
            if (k >= 1 && k <= 32) threads = (unsigned)k;
        }
        size_t slot_bytes = kStageSlotBytes;
        if (const char* kb = getenv("EZK_STAGE_SLOT_KB")) {  // tests: small slots => many chunks per column
            const long k = atol(kb);
            if (k >= 4 && k <= (long)(kStageSlotBytes >> 10)) slot_bytes = (size_t)k << 10;
        }
        // the ring and its events first, the pool last: a failed allocation leaves the prover on the plain copy
        // instead of half-initialised (later proofs would otherwise copy into null slots)
        for (int i = 0; i < kStageSlots; i++) {
            if (!stage_[i] && cudaMallocHost(&stage_[i], kStageSlotBytes) != cudaSuccess) {
                cudaGetLastError();
                stage_[i] = nullptr;
                return false;
            }
            if (!stage_ev_[i] && cudaEventCreateWithFlags(&stage_ev_[i], cudaEventDisableTiming) != cudaSuccess) {
                cudaGetLastError();
                stage_ev_[i] = nullptr;
                return false;
            }
        }
        stage_slot_bytes_ = slot_bytes;
        copy_pool_.reset(new CopyPool(threads));
    }
    return true;
}

// One column through the ring: slot k is refilled once the copy that last read it has finished.
void GpuProver::staged_copy_column(uint4* d_dst, const uint8_t* src, size_t bytes, uint64_t& chunk) {
    for (size_t off = 0; off < bytes; off += stage_slot_bytes_, chunk++) {
        const size_t len = std::min(stage_slot_bytes_, bytes - off);
        const int slot = (int)(chunk % kStageSlots);
        EZK_CUDA(cudaEventSynchronize(stage_ev_[slot]));  // returns at once for an event never recorded
        copy_pool_->copy(stage_[slot], src + off, len);
        EZK_CUDA(cudaMemcpyAsync((uint8_t*)d_dst + off, stage_[slot], len, cudaMemcpyHostToDevice, copy_stream_));
        EZK_CUDA(cudaEventRecord(stage_ev_[slot], copy_stream_));
    }
}

void GpuProver::join(int rank, int world, const uint8_t id[128]) {
    EZK_CUDA(cudaSetDevice(device_));
    sync();
    comm_.init(rank, world, id);
}

void GpuProver::join_local(int rank, LocalGroup* group) {
    EZK_CUDA(cudaSetDevice(device_));
    sync();
    comm_.init_local(rank, group);
}

void GpuProver::timer_start() {
    EZK_CUDA(cudaSetDevice(device_));
    EZK_CUDA(cudaEventRecord(timer_ev_[0], stream_));
}
float GpuProver::timer_stop() {
    EZK_CUDA(cudaSetDevice(device_));
    EZK_CUDA(cudaEventRecord(timer_ev_[1], stream_));
    EZK_CUDA(cudaEventSynchronize(timer_ev_[1]));
    float ms = 0;
    EZK_CUDA(cudaEventElapsedTime(&ms, timer_ev_[0], timer_ev_[1]));
    return ms;
}

GpuProver::~GpuProver() {
    cudaSetDevice(device_);
    for (auto& e : ev_) cudaEventDestroy(e);
    for (auto& e : timer_ev_) cudaEventDestroy(e);
    for (auto& e : copy_ev_) cudaEventDestroy(e);
    for (auto& e : col_ev_) cudaEventDestroy(e);
    for (auto& e : grp_ev_) cudaEventDestroy(e);
    for (auto& e : share_ev_) cudaEventDestroy(e);
    cudaStreamDestroy(copy_stream_);
    for (auto& e : aux_ev_) cudaEventDestroy(e);
    cudaStreamDestroy(aux_stream_);
    cudaFree(d_flag_);
    cudaFree(d_params_);
    cudaFree(d_bden_);
    cudaFreeHost(pinned_);
    for (int i = 0; i < kStageSlots; i++) {
        if (stage_ev_[i]) cudaEventDestroy(stage_ev_[i]);
        if (stage_[i]) cudaFreeHost(stage_[i]);
    }
    cudaFree(arena_.base);
    ntt_tables_free(tables_);
    cudaStreamDestroy(stream_);
}

void GpuProver::reserve(size_t elems) {
    if (elems <= arena_.capacity) return;
    if (arena_.base) {
        EZK_CUDA(cudaStreamSynchronize(stream_));
        EZK_CUDA(cudaFree(arena_.base));
        arena_.base = nullptr, arena_.capacity = 0;
    }
    last_ = Last{};
    EZK_CUDA(cudaMalloc(&arena_.base, elems * sizeof(uint4)));
    arena_.capacity = elems;
    arena_generation_ = ((uint64_t)getpid() << 32) | (uint64_t)(++arena_allocations_);
}

uint4* GpuProver::alloc(size_t elems) {
    elems = (elems + 15) & ~(size_t)15;  // 256-byte granules
    if (arena_.used + elems > arena_.capacity) throw ProveFailure{EZK_ERR_INTERNAL, "device workspace exhausted"};
    uint4* p = arena_.base + arena_.used;
    arena_.used += elems;
    return p;
}

void GpuProver::sync() { EZK_CUDA(cudaStreamSynchronize(stream_)); }

std::vector<uint8_t> GpuProver::prove(const uint8_t* const* host_columns, const uint4* device_trace, uint64_t n,
                                      const PublicInputs& pub, const ProofOptions& opt, const OpList* ops) {
    try {
        return prove_impl(host_columns, device_trace, n, pub, opt, ops);
    } catch (...) {
        comm_.abort();  // in-process group: release the peers waiting in a collective
        throw;
    }
}

std::vector<uint8_t> GpuProver::prove_impl(const uint8_t* const* host_columns, const uint4* device_trace, uint64_t n,
                                           const PublicInputs& pub, const ProofOptions& opt, const OpList* ops) {
    if (ops && (!host_columns || !ops->codes || !ops->last_row || ops->count >= n))
        throw ProveFailure{EZK_ERR_INVALID_ARGUMENT, "operation list: needs host columns, codes, the last row and count < trace length"};
    if (opt.field_ext != 1) throw ProveFailure{EZK_ERR_UNSUPPORTED_FIELD_EXTENSION, "only FieldExtension::None is supported"};
    if (opt.blowup != 8 || opt.fri_fold != 8)
        throw ProveFailure{EZK_ERR_INVALID_ARGUMENT, "blowup factor and FRI folding factor must both be 8"};
    if (n < 64 || (n & (n - 1)) || n > (1ull << 24))
        throw ProveFailure{EZK_ERR_INVALID_ARGUMENT, "trace length must be a power of two in [2^6, 2^24]"};
    if (pub.lwe_k != 4) throw ProveFailure{EZK_ERR_INVALID_ARGUMENT, "the AIR requires lwe_k = 4 (lwe_size 5)"};
    if (opt.num_queries == 0 || opt.num_queries > 255 || opt.grinding > 32 || opt.fri_rem_max_deg > 255 ||
        ((opt.fri_rem_max_deg + 1) & opt.fri_rem_max_deg))
        throw ProveFailure{EZK_ERR_INVALID_ARGUMENT, "unsupported proof options"};
    EZK_CUDA(cudaSetDevice(device_));

    const uint64_t L = n * 8;
    const uint32_t log_n = ilog2_u64(n), log_L = log_n + 3;
    const Fp o = Fp::from_u64(kDomainOffset), o_inv = inverse(o);
    const Fp g = root_of_unity(log_n);
    const size_t nlayers = num_fri_layers(L, opt);
    const uint32_t eval_blocks = log_n > 14 ? 1u << (log_n - 14) : 1;

    // ---- multi-GPU layout.  Rank r of G owns the LDE rows i = r mod G (whole cosets: the (i, i + 8) frames
    // of the constraint evaluation and the 8 points of a FRI row stay on one GPU), interpolates the trace columns
    // c = r mod G, and holds the Merkle subtree over the leaves [r L / G, (r + 1) L / G) of every commitment.  The
    // same code runs with G = 1 when EZK_FORCE_SHARDED_PATH=1 (tests: every collective becomes a local copy).
    const uint32_t G = (uint32_t)comm_.world(), glog = comm_.world_log(), me = (uint32_t)comm_.rank();
    bool sharded = comm_.active();
    if (const char* f = getenv("EZK_FORCE_SHARDED_PATH")) sharded = sharded || f[0] == '1';
    const RowShard sh{me, glog};
    const CosetSet cs{3 - glog, me, G};
    const uint32_t cn = 8 / G;                      // cosets per rank
    const uint64_t L_local = L >> glog;
    const uint32_t rounds = (kWidth + G - 1) / G;   // trace columns per rank (the last round may be partial)
    uint64_t shard_fri_min_rows = 4096;             // FRI layers with fewer rows are gathered and finished on every rank
    if (const char* e = getenv("EZK_SHARD_FRI_MIN_ROWS")) shard_fri_min_rows = std::max<uint64_t>(64, (uint64_t)atoll(e));

    // trace columns per interpolation + extension launch of a device-resident trace: the transform scratch is sized by
    // the group, which is what lets a 2^24-row proof fit one GPU (EZK_LDE_GROUP overrides)
    uint32_t lde_group = log_n >= 23 ? 4 : kWidth;
    if (const char* e = getenv("EZK_LDE_GROUP")) lde_group = (uint32_t)std::min<long>(kWidth, std::max<long>(1, atol(e)));

    // ---- workspace (16-byte elements; the arena is sized once per trace length and kept) ----
    // transform scratch in columns of L_local elements: sharded batches hold <= 8 columns, the host path 2-column groups,
    // the DEEP step 2 columns; the composition columns are extended in groups that fit
    // columns per launch of a host trace: groups grow from 1 to this cap while the upload runs ahead (see below)
    uint32_t host_group_cap = std::min(8u, lde_group);
    if (const char* e = getenv("EZK_HOST_GROUP_CAP")) host_group_cap = (uint32_t)std::min<long>(lde_group, std::max<long>(1, atol(e)));
    const uint32_t tmp_cols = sharded ? 8 : std::max(2u, host_columns ? host_group_cap : lde_group);
    const uint32_t comp_group = std::min(kCompCols, tmp_cols);
    size_t fri_elems = 0;
    uint64_t first_whole_layer = 0;  // sharded: size of the first FRI layer that is gathered onto every rank
    {
        uint64_t s = L;
        bool packed = sharded;
        for (size_t k = 0; k < nlayers; k++) {
            const uint64_t m = s / 8;
            if (packed && !(m >= shard_fri_min_rows && m >= (uint64_t)G * G)) packed = false, first_whole_layer = s;
            fri_elems += s / 8 + 4 * (s / 8) + 64 + (first_whole_layer == s ? s : 0);
            s /= 8;
        }
        if (packed) first_whole_layer = s, fri_elems += s;
        fri_elems += 2 * s + 8192;
    }
    size_t fri_tree_elems = 0;  // subtrees of the sharded FRI layers
    {
        uint64_t s = L;
        for (size_t k = 0; k < nlayers && sharded; k++, s /= 8) {
            const uint64_t m = s / 8;
            if (!(m >= shard_fri_min_rows && m >= (uint64_t)G * G)) break;
            fri_tree_elems += 4 * (m >> glog) + 16;
        }
    }
    const size_t tcoef_cols = sharded ? (size_t)rounds * G : kWidth;
    const size_t allg_elems = sharded ? std::max<size_t>(first_whole_layer, (size_t)32768 * G) + 64 : 0;
    const size_t need = (host_columns ? (size_t)kWidth * n : 0) + tcoef_cols * n + (size_t)kWidth * L_local + (size_t)tmp_cols * L_local +
                        (sharded ? 0 : 8 * L) /* whole trees */ + 3 * L_local + L + (size_t)kCompCols * L_local + 2 * n +
                        (sharded ? 0 : 2 * L) + (size_t)(2 * kWidth + kCompCols) * eval_blocks + 4096 +
                        (sharded ? 4 * L_local + (size_t)cn * n + 8 * n + allg_elems + 8 * L_local /* subtrees */ : 0) + fri_elems +
                        8192 + 32768 + 64 * 32 + (ops ? n / 16 + n / 4096 + 128 : 0) + fri_tree_elems + 256;
    reserve(need);
    reset_arena();
    // Peer-visible part of the workspace first, so that it sits at the same offset on every rank whatever the input mode:
    // the Merkle subtrees, whose leaves the other ranks' row-hash kernels write directly over NVLink (fused exchange)
    uint4* d_peer_scratch = sharded ? alloc(128) : nullptr;  // handle exchange (first KiB), barrier words (second)
    uint4* d_trees = sharded ? alloc(8 * L_local + fri_tree_elems) : nullptr;
    uint4* d_recv = sharded ? alloc(2 * L_local) : nullptr;  // digests of this rank's leaf range, one chunk per sender
    size_t trees_used = 0;
    bool peer_stores = false;
    if (sharded && G > 1) {
        const char* e = getenv("EZK_PEER_STORES");
        const bool mapped = comm_.map_peers(arena_.base, arena_.capacity * sizeof(uint4), arena_generation_, pinned_, d_peer_scratch, stream_);
        peer_stores = mapped && !(e && e[0] == '0');
    }
    uint4* d_trace_in = host_columns ? alloc(kWidth * n) : nullptr;
    uint4* d_tcoef = alloc(tcoef_cols * n);
    uint4* d_tlde = alloc(kWidth * L_local);  // multi-GPU: this rank's rows only, packed order
    uint4* d_tmp = alloc((size_t)tmp_cols * L_local);
    uint4* d_tnodes = sharded ? nullptr : alloc(4 * L);
    uint4* d_invden = alloc(L_local);
    uint4* d_combined = alloc(L_local);
    uint4* d_ccoef = alloc(L);
    uint4* d_clde = alloc(kCompCols * L_local);
    uint4* d_cnodes = sharded ? nullptr : alloc(4 * L);
    uint4* d_pq = alloc(2 * n);
    uint4* d_pqlde = sharded ? nullptr : alloc(2 * L);
    uint4* d_deep = alloc(L_local);
    uint4* d_scratch = alloc((size_t)(2 * kWidth + kCompCols) * eval_blocks);
    uint4* d_small = alloc(4096);  // OOD outputs, deep coefficients, subtree roots, flags
    uint8_t* d_codes = ops ? reinterpret_cast<uint8_t*>(alloc(n / 16 + 1)) : nullptr;
    uint32_t* d_scan = ops ? reinterpret_cast<uint32_t*>(alloc(n / 4096 + 64)) : nullptr;
    // sharded: packed per-row products of this rank, the receive side of the exchanges
    uint4* d_pack = sharded ? alloc(2 * L_local) : nullptr;
    uint4* d_rloc = sharded ? alloc((size_t)cn * n) : nullptr;   // per-coset interpolations of the constraint evaluations
    uint4* d_rall = sharded ? alloc(8 * n) : nullptr;
    uint4* d_allg = sharded ? alloc(allg_elems) : nullptr;       // gathered FRI evaluations / opened rows of every rank

    struct ShardTree {            // Merkle tree of one commitment
        bool split = false;       // false: `nodes` is the whole tree (heap order, node 1 = root)
        uint4* nodes = nullptr;   // split: this rank's subtree over leaves_local leaves (node 1 = subtree root)
        uint64_t leaves = 0, leaves_local = 0;
        std::vector<Hash32> top;  // split: heap order, [G, 2G) = subtree roots of the ranks, [1, G) the levels above
    };

    last_ = Last{};
    last_.n = n, last_.L = L, last_.tcoef = d_tcoef;
    if (!sharded) last_.tlde = d_tlde, last_.clde = d_clde, last_.combined_copy = d_combined, last_.deep = d_deep;

    int evi = 0;
    bool mark_pending = false;
    auto mark = [&]() { EZK_CUDA(cudaEventRecord(ev_[evi++], stream_)); };
    auto d2h = [&](void* dst, const void* src, size_t bytes) {
        if (bytes > pinned_bytes_) throw ProveFailure{EZK_ERR_INTERNAL, "staging buffer too small"};
        EZK_CUDA(cudaMemcpyAsync(pinned_, src, bytes, cudaMemcpyDeviceToHost, stream_));
        sync();
        memcpy(dst, pinned_, bytes);
    };
    auto h2d = [&](void* dst, const void* src, size_t bytes) {
        if (bytes > pinned_bytes_) throw ProveFailure{EZK_ERR_INTERNAL, "staging buffer too small"};
        sync();  // the staging buffer may still be in flight
        memcpy(pinned_, src, bytes);
        EZK_CUDA(cudaMemcpyAsync(dst, pinned_, bytes, cudaMemcpyHostToDevice, stream_));
    };
    // commitment to a column-major table of `leaves` rows (LDE tables, FRI layers; sharded: this rank's packed rows)
    auto commit = [&](const uint4* table, uint32_t width, uint64_t leaves, uint4* whole_nodes, ShardTree& tree) -> Hash32 {
        Hash32 root;
        tree.leaves = leaves;
        if (!sharded) {
            tree.split = false, tree.nodes = whole_nodes, tree.leaves_local = leaves;
            merkle_hash_rows(stream_, table, leaves, width, leaves, whole_nodes);
            merkle_build(stream_, whole_nodes, leaves);
            d2h(root.data(), whole_nodes + 2, 32);  // node 1
            return root;
        }
        const uint64_t ll = leaves >> glog;  // leaves of this rank's subtree
        // rank r's packed digests are those of rows r, r + G, ...: chunk q of them lies in rank q's leaf range
        tree.split = true, tree.leaves_local = ll, tree.nodes = d_trees + trees_used;
        trees_used += (4 * ll + 15) & ~(size_t)15;
        if (trees_used > 8 * L_local + fri_tree_elems) throw ProveFailure{EZK_ERR_INTERNAL, "subtree workspace exhausted"};
        const uint64_t chunk = ll >> glog;  // rows per destination
        if (peer_stores) {
            // exchange fused into the row hash: every digest goes straight into the receive area of the rank that owns
            // its subtree (NVLink peer stores), where an all-to-all would have put it
            uint4* peer_recv[8] = {nullptr};
            for (uint32_t q = 0; q < G; q++) peer_recv[q] = comm_.peer((int)q, d_recv);
            hash_rows_to_peers(stream_, table, ll, width, ll, sh, peer_recv);
            comm_.barrier(d_peer_scratch + 64, stream_);  // all ranks' digests have landed
            count_launch();
        } else {
            hash_rows_sharded(stream_, table, ll, width, ll, sh, d_pack);
            comm_.all_to_all(d_pack, d_recv, chunk * 32, stream_);
            count_launch();
        }
        // (measured and dropped: hashing and exchanging destination by destination in a ring-shifted pipeline of NCCL
        // send/recv pairs - 7 small steps cost more than the one all-to-all they hide: trace commitment 2.2 -> 3.2 ms
        // at 2^22 rows on 8 GPUs)
        unpack_rows(stream_, d_recv, ll >> glog, glog, 2, tree.nodes + 2 * ll);
        merkle_build(stream_, tree.nodes, ll);
        uint4* d_roots = d_small + 1024;
        comm_.all_gather(tree.nodes + 2, d_roots, 32, stream_);  // the G subtree roots
        count_launch();
        std::vector<Hash32> roots(G);
        d2h(roots.data(), d_roots, 32 * G);
        tree.top.assign(2 * G, Hash32{});
        for (uint32_t q = 0; q < G; q++) tree.top[G + q] = roots[q];
        for (uint32_t k = G - 1; k >= 1; k--) tree.top[k] = merge_digests(tree.top[2 * k], tree.top[2 * k + 1]);
        return tree.top[1];
    };

    // ---- (0) transcript ----
    RandomCoin coin;
    coin.init(coin_seed(kWidth, n, opt, pub.elements));
    std::vector<uint8_t> commitments;

    // ---- (1) trace upload, interpolation, LDE, commitment ----
    mark();
    EZK_CUDA(cudaMemsetAsync(d_flag_, 0, sizeof(uint32_t), stream_));
    auto uploaded = [&](uint32_t c) { return !(ops && is_op_column(c)); };  // columns that come from the caller's memory
    if (ops) {
        // clk, op bits, chiplet flag and stack depth from the operation list, on the device (trace/expand.cuh)
        EZK_CUDA(cudaMemcpyAsync(d_codes, ops->codes, ops->count, cudaMemcpyHostToDevice, stream_));
        h2d(d_small + 3072, ops->last_row, kWidth * 16);
        expand_op_columns(stream_, d_codes, ops->count, n, pub.lwe_k + 1, d_small + 3072, d_scan, d_trace_in, d_flag_);
    }
    {
        NttScale sc{};
        put(sc.cvec[0], inverse(Fp::from_u64(n)));
        sc.chunk_shift = 63, sc.use_offset = 1;
        if (sharded) {
            // this rank interpolates the columns c = me, me + G, ... (uploading only those); round j of every rank
            // (columns [jG, jG + G)) is all-gathered in place on the auxiliary stream, and the extension of a batch of
            // rounds (about 8 columns per launch) over this rank's cosets starts as soon as the batch has arrived
            EZK_CUDA(cudaEventRecord(copy_ev_[15], stream_));
            EZK_CUDA(cudaStreamWaitEvent(copy_stream_, copy_ev_[15], 0));  // the arena may still be in use
            EZK_CUDA(cudaStreamWaitEvent(aux_stream_, copy_ev_[15], 0));
            const bool staged = host_columns && use_staged_upload(host_columns);
            const uint4* own_src = device_trace ? device_trace : d_trace_in;
            const uint32_t rb = std::max(1u, 8u / G);  // rounds per batch: about 8 columns per extension launch
            uint64_t chunk = 0;
            mark_pending = true;
            auto extend_batch = [&](uint32_t j0) {
                const uint32_t j1 = std::min(rounds, j0 + rb);
                EZK_CUDA(cudaStreamWaitEvent(stream_, share_ev_[2 * (j1 - 1) + 1], 0));
                const size_t c0 = (size_t)j0 * G;
                const uint32_t cols = std::min<uint32_t>((j1 - j0) * G, kWidth - (uint32_t)c0);
                lde_columns(tables_, stream_, d_tcoef + c0 * n, n, d_tlde + c0 * L_local, L_local, d_tmp, cols, log_n, cs);
            };
            if (device_trace) {
                // resident trace: interpolate all own columns at once, start every all-gather, and extend the OWN
                // columns first - they need no exchange, so the first all-gathers travel under useful work; then the
                // columns of the other ranks, batch by batch as they arrive
                const uint32_t mine = (kWidth - me + G - 1) / G;
                mark(), mark_pending = false;
                for (uint32_t k = 0; k < mine; k++) check_canonical(stream_, own_src + (size_t)(k * G + me) * n, n, d_flag_);
                ntt_columns(tables_, stream_, own_src + (size_t)me * n, (uint64_t)G * n, d_tcoef + (size_t)me * n, (uint64_t)G * n, d_tmp,
                            mine, log_n, true, &sc);
                EZK_CUDA(cudaEventRecord(share_ev_[0], stream_));
                EZK_CUDA(cudaStreamWaitEvent(aux_stream_, share_ev_[0], 0));
                for (uint32_t j = 0; j < rounds; j++) {
                    comm_.all_gather(d_tcoef + (size_t)(j * G + me) * n, d_tcoef + (size_t)j * G * n, n * 16, aux_stream_);
                    count_launch();
                    EZK_CUDA(cudaEventRecord(share_ev_[2 * j + 1], aux_stream_));
                }
                // columns c0, c0 + step, ... (count of them) in one launch, at most 8 at a time (scratch)
                auto extend_progression = [&](uint32_t c0, uint32_t step, uint32_t count) {
                    for (uint32_t k = 0; k < count; k += 8)
                        lde_columns(tables_, stream_, d_tcoef + (size_t)(c0 + k * step) * n, (uint64_t)step * n,
                                    d_tlde + (size_t)(c0 + k * step) * L_local, (uint64_t)step * L_local, d_tmp, std::min(8u, count - k),
                                    log_n, cs);
                };
                extend_progression(me, G, mine);
                for (uint32_t j0 = 0; j0 < rounds; j0 += rb) {
                    const uint32_t j1 = std::min(rounds, j0 + rb);
                    EZK_CUDA(cudaStreamWaitEvent(stream_, share_ev_[2 * (j1 - 1) + 1], 0));
                    if (G - 1 <= 2 * (j1 - j0)) {
                        // one launch per foreign rank q: its columns j G + q of the batch
                        for (uint32_t q = 0; q < G; q++) {
                            if (q == me) continue;
                            uint32_t count = 0;
                            while (j0 + count < j1 && (j0 + count) * G + q < kWidth) count++;
                            if (count) extend_progression(j0 * G + q, G, count);
                        }
                    } else {
                        // per round the two contiguous ranges on either side of the own column
                        for (uint32_t j = j0; j < j1; j++) {
                            const uint32_t lo = j * G, hi = std::min(kWidth, lo + G), own = lo + me;
                            if (own > lo) extend_progression(lo, 1, std::min(own, hi) - lo);
                            if (own + 1 < hi) extend_progression(own + 1, 1, hi - own - 1);
                        }
                    }
                }
            } else
            // software pipeline on the compute stream: interpolate(b), extend(b - 1), interpolate(b + 1), ... so that the
            // all-gather of batch b (auxiliary stream) and the upload of batch b + 1 run under the extension of batch b - 1
            {
            for (uint32_t j0 = 0; j0 < rounds; j0 += rb) {
                const uint32_t j1 = std::min(rounds, j0 + rb);
                uint32_t own = 0;  // own columns of this batch: c = (j0 + k) G + me < 28
                for (uint32_t j = j0; j < j1; j++) {
                    const uint32_t c = j * G + me;
                    if (c >= kWidth) break;
                    own++;
                    if (host_columns && uploaded(c)) {
                        if (staged)
                            staged_copy_column(d_trace_in + (size_t)c * n, host_columns[c], n * 16, chunk);
                        else
                            EZK_CUDA(cudaMemcpyAsync(d_trace_in + (size_t)c * n, host_columns[c], n * 16, cudaMemcpyHostToDevice, copy_stream_));
                    }
                }
                if (host_columns) {
                    EZK_CUDA(cudaEventRecord(copy_ev_[(j0 / rb) % 14], copy_stream_));
                    EZK_CUDA(cudaStreamWaitEvent(stream_, copy_ev_[(j0 / rb) % 14], 0));
                }
                if (mark_pending) mark(), mark_pending = false;
                if (own) {
                    const size_t first = (size_t)(j0 * G + me) * n;
                    for (uint32_t k = 0; k < own; k++) check_canonical(stream_, own_src + first + (size_t)k * G * n, n, d_flag_);
                    ntt_columns(tables_, stream_, own_src + first, (uint64_t)G * n, d_tcoef + first, (uint64_t)G * n, d_tmp, own, log_n, true, &sc);
                }
                EZK_CUDA(cudaEventRecord(share_ev_[2 * j0], stream_));
                EZK_CUDA(cudaStreamWaitEvent(aux_stream_, share_ev_[2 * j0], 0));
                for (uint32_t j = j0; j < j1; j++) {
                    comm_.all_gather(d_tcoef + (size_t)(j * G + me) * n, d_tcoef + (size_t)j * G * n, n * 16, aux_stream_);
                    count_launch();
                    EZK_CUDA(cudaEventRecord(share_ev_[2 * j + 1], aux_stream_));
                }
                if (j0 >= rb) extend_batch(j0 - rb);
            }
            extend_batch(((rounds - 1) / rb) * rb);
            }
        } else if (host_columns) {
            // Upload and transform overlap: columns travel on the copy stream while the interpolation + LDE of earlier
            // columns run on the compute stream (columns are independent until the row hash).  How many columns go into
            // one launch is decided as they arrive (host/launch_groups.h): the groups grow 1, 1, 2, 2, 3, ... up to
            // host_group_cap columns while the upload runs ahead, and the GPU never waits for a group to fill.
            EZK_CUDA(cudaEventRecord(copy_ev_[15], stream_));
            EZK_CUDA(cudaStreamWaitEvent(copy_stream_, copy_ev_[15], 0));  // the arena may still be in use
            const bool staged = use_staged_upload(host_columns);
            uint64_t chunk = 0;
            if (!staged)
                for (uint32_t c = 0; c < kWidth; c++)
                    if (uploaded(c)) {
                        EZK_CUDA(cudaMemcpyAsync(d_trace_in + (size_t)c * n, host_columns[c], n * 16, cudaMemcpyHostToDevice,
                                                 copy_stream_));
                        EZK_CUDA(cudaEventRecord(col_ev_[c], copy_stream_));
                    }
            uint32_t launched = 0, ngroups = 0, head = 0;  // groups [head, ngroups) may still be running
            uint32_t group_cols[32];
            auto launch = [&](uint32_t c1) {  // transforms of columns [launched, c1), behind the upload of the last of them
                for (uint32_t c = c1; c-- > launched;)
                    if (uploaded(c)) {
                        EZK_CUDA(cudaStreamWaitEvent(stream_, col_ev_[c], 0));  // the copy stream is in order
                        break;
                    }
                if (launched == 0) mark();
                const uint32_t cols = c1 - launched;
                const size_t c0 = launched;
                check_canonical(stream_, d_trace_in + c0 * n, (size_t)cols * n, d_flag_);
                ntt_columns(tables_, stream_, d_trace_in + c0 * n, n, d_tcoef + c0 * n, n, d_tmp, cols, log_n, true, &sc);
                lde_columns(tables_, stream_, d_tcoef + c0 * n, n, d_tlde + c0 * L, L, d_tmp, cols, log_n, cs);
                EZK_CUDA(cudaEventRecord(grp_ev_[ngroups], stream_));
                group_cols[ngroups++] = cols;
                launched = c1;
            };
            auto should_launch = [&](uint32_t avail, bool all_sent) {
                if (avail == 0) return false;
                while (head < ngroups && cudaEventQuery(grp_ev_[head]) == cudaSuccess) head++;
                cudaGetLastError();  // cudaErrorNotReady is not an error
                const uint32_t pending = ngroups - head;
                return host_group_ready(avail, pending, pending ? group_cols[head] : 0, host_group_cap, all_sent);
            };
            if (staged) {  // this thread feeds the ring: decide after every column it has sent
                for (uint32_t c = 0; c < kWidth; c++) {
                    if (uploaded(c)) {
                        staged_copy_column(d_trace_in + (size_t)c * n, host_columns[c], n * 16, chunk);
                        EZK_CUDA(cudaEventRecord(col_ev_[c], copy_stream_));
                    }
                    if (should_launch(c + 1 - launched, false)) launch(std::min(c + 1, launched + host_group_cap));
                }
                while (launched < kWidth) launch(std::min(kWidth, launched + host_group_cap));
            } else {  // all copies are queued: follow their completion
                uint32_t ready = 0;
                while (launched < kWidth) {
                    while (ready < kWidth && (!uploaded(ready) || cudaEventQuery(col_ev_[ready]) == cudaSuccess)) ready++;
                    cudaGetLastError();
                    if (should_launch(ready - launched, ready == kWidth))
                        launch(std::min(ready, launched + host_group_cap));
                    else if (ready == launched)
                        EZK_CUDA(cudaEventSynchronize(col_ev_[ready]));  // nothing to do until the next column is here
                }
            }
        } else {
            mark();
            check_canonical(stream_, device_trace, kWidth * n, d_flag_);
            // column groups: the coefficient columns of a group (16 MiB each at 2^20) and the inter-pass twiddle table
            // compete for the L2, and the transform scratch is sized by the group
            for (uint32_t c0 = 0; c0 < kWidth; c0 += lde_group) {
                const uint32_t cols = std::min(lde_group, kWidth - c0);
                ntt_columns(tables_, stream_, device_trace + (size_t)c0 * n, n, d_tcoef + (size_t)c0 * n, n, d_tmp, cols, log_n, true, &sc);
                lde_columns(tables_, stream_, d_tcoef + (size_t)c0 * n, n, d_tlde + (size_t)c0 * L, L, d_tmp, cols, log_n, cs);
            }
        }
    }
    mark();
    // the divisor inverses of the boundary constraints depend on n (and the row shard) only: cached across proofs of
    // the same length; the first proof computes them on the auxiliary stream while the (latency-bound) Merkle tree
    // of the trace commitment is built and its root travels to the host
    const Fp g_last = pow(g, n - 2), g_last2 = pow(g, n - 1);
    const uint4* d_bden = d_invden;
    {
        const uint64_t key = ((uint64_t)log_n << 16) | ((uint64_t)sh.world_log << 8) | sh.rank | (1ull << 40);
        bool have = d_bden_ && bden_key_ == key;
        if (!have) {
            if (d_bden_) cudaFree(d_bden_);  // no proof is in flight: prove() ends with a stream synchronisation
            d_bden_ = nullptr, bden_key_ = 0;
            if (cudaMalloc(&d_bden_, L_local * sizeof(uint4)) != cudaSuccess) {
                cudaGetLastError();
                d_bden_ = nullptr;
            }
        }
        uint4* target = d_bden_ ? d_bden_ : d_invden;
        EZK_CUDA(cudaEventRecord(aux_ev_[0], stream_));
        EZK_CUDA(cudaStreamWaitEvent(aux_stream_, aux_ev_[0], 0));
        if (!have) {
            uint64_t a[2], b[2];
            put(a, Fp(1)), put(b, g_last);
            domain_pair_inverse(aux_stream_, tables_.root_fwd, tables_.root_inv, log_L, a, b, target, sh);
            if (d_bden_) bden_key_ = key;
        }
        EZK_CUDA(cudaEventRecord(aux_ev_[1], aux_stream_));
        d_bden = target;
    }
    ShardTree trace_tree, comp_tree;
    last_.trace_root = commit(d_tlde, kWidth, L, d_tnodes, trace_tree);
    commitments.insert(commitments.end(), last_.trace_root.begin(), last_.trace_root.end());
    coin.reseed(last_.trace_root);
    mark();

    // ---- (2) constraint evaluation ----
    {
        ConstraintParams hp{};
        for (uint32_t j = 0; j < kTransitions; j++) {
            const Fp c = coin.draw();
            put(hp.tcoef[j], c), put_pre(hp.tcoef_pre[j], c);
        }
        Fp bc[kAssertions];
        for (uint32_t k = 0; k < kAssertions; k++) bc[k] = coin.draw();
        // assertions sorted by (step, column): air/src/lib.rs:170-195 + winter-air's prepare_assertions
        const uint32_t cols0[12] = {0, 7, 8, 11, 12, 13, 14, 15, 16, 17, 18, 19};
        const uint32_t cols1[10] = {7, 8, 12, 13, 14, 15, 16, 17, 18, 19};
        for (uint32_t k = 0; k < 12; k++) {
            put(hp.bcoef[k], bc[k]), put_pre(hp.bcoef_pre[k], bc[k]);
            put(hp.bval[k], Fp());
            hp.bcol[k] = cols0[k];
        }
        for (uint32_t k = 0; k < 10; k++) {
            put(hp.bcoef[12 + k], bc[12 + k]), put_pre(hp.bcoef_pre[12 + k], bc[12 + k]);
            // values: program_hash[0..2] for columns 7,8; stack_outputs[0..8] for columns 12..19
            put(hp.bval[12 + k], k < 2 ? pub.elements[k] : pub.elements[2 + (k - 2)]);
            hp.bcol[12 + k] = cols1[k];
        }
        {
            Fp c1;
            for (uint32_t k = 0; k < 10; k++) c1 = c1 + bc[12 + k] * (k < 2 ? pub.elements[k] : pub.elements[2 + (k - 2)]);
            put(hp.bsum1, c1);
        }
        hp.delta = pub.lwe_delta;
        // 1/(x^n - 1): x_i^n = o^n * w_8^(i mod 8)
        const Fp on = pow(o, n), w8 = root_of_unity(3);
        Fp w8c(1);
        for (int c = 0; c < 8; c++) {
            put(hp.inv_zn[c], inverse(on * w8c - Fp(1)));
            w8c = w8c * w8;
        }
        put(hp.g_last, g_last), put(hp.g_last2, g_last2);
        for (int i = 0; i < 16; i++) put(hp.inv_mds[i], rescue_const(rescue_inv_mds()[i])), put_pre(hp.inv_mds_pre[i], rescue_const(rescue_inv_mds()[i]));
        // periodic columns: mask + 8 ARK columns (air/src/lib.rs:201-225, rescue.rs:120-134), interpolated over
        // <w_16> and tabulated at x^(n/16) for the 128 distinct values of step mod 128
        // (depends on n only: cached per trace length, 1152 Horner evaluations otherwise)
        if (ptable_log_n_ != log_n) {
            const std::vector<std::vector<Fp>> polys = periodic_polys();
            const Fp on16 = pow(o, n / kCycle), w128 = root_of_unity(7);
            Fp wr(1);
            ptable_.assign(128 * kPeriodic, Fp());
            for (uint32_t r = 0; r < 128; r++) {
                const Fp y = on16 * wr;
                for (uint32_t p = 0; p < kPeriodic; p++) ptable_[r * kPeriodic + p] = horner(polys[p], y);
                wr = wr * w128;
            }
            ptable_log_n_ = log_n;
        }
        for (uint32_t k = 0; k < 128 * kPeriodic; k++) put(hp.ptable[k], ptable_[k]);
        h2d(d_params_, &hp, sizeof(hp));
        EZK_CUDA(cudaStreamWaitEvent(stream_, aux_ev_[1], 0));
        // sharded: one contiguous array per owned coset, which is what the per-coset interpolation below reads
        evaluate_constraints(stream_, tables_.root_fwd, d_tlde, L_local, log_L, d_params_, d_bden, d_combined, sh, sharded);
    }
    mark();

    // ---- (3) composition polynomial: interpolate over the coset, 7 columns of n, LDE, commitment ----
    {
        const Fp l_inv = inverse(Fp::from_u64(L)), o_n_inv = pow(o_inv, n);
        uint64_t scale8[8][2];
        Fp acc = l_inv;
        for (int j = 0; j < 8; j++) {
            put(scale8[j], acc);  // 3^(-jn) / L
            acc = acc * o_n_inv;
        }
        uint32_t flag = 0;
        if (sharded) {
            // C(x) = sum_j x^(jn) H_j(x) restricted to coset c is R_c(x) = sum_j (3^n w_8^c)^j H_j(x), degree < n: every
            // rank interpolates its cosets; an all-to-all hands rank q the coefficient slice [q n / G, (q + 1) n / G) of
            // all 8 cosets, where an 8-point inverse DFT across the cosets separates the H_j (composition_recombine);
            // the slices of column j are then all-gathered (auxiliary stream) while column j - 1 is being extended
            const uint64_t slice = n >> glog;
            ntt_columns(tables_, stream_, d_combined, n, d_rloc, n, d_tmp, cn, log_n, true, nullptr);
            for (uint32_t k = 0; k < cn; k++) {  // coset me + G k of this rank -> array me + G k of every rank's slice set
                comm_.all_to_all(d_rloc + (size_t)k * n, d_rall + (size_t)k * n, slice * 16, stream_);
                count_launch();
            }
            // d_rall[(k G + q) * slice ..] = coset q + G k: natural coset order, one slice each; d_rloc is free again
            composition_recombine(tables_, stream_, d_rall, log_n, (uint64_t)me * slice, slice, scale8, d_rloc, d_flag_);
            uint32_t* d_flags = reinterpret_cast<uint32_t*>(d_small + 2048);
            comm_.all_gather(d_flag_, d_flags, sizeof(uint32_t), stream_);  // every rank takes the same decision
            count_launch();
            uint32_t flags[8] = {0};
            d2h(flags, d_flags, sizeof(uint32_t) * G);
            for (uint32_t q = 0; q < G; q++) flag |= flags[q];
        } else {
            NttScale sc{};
            memcpy(sc.cvec, scale8, sizeof(scale8));
            sc.chunk_shift = log_n, sc.use_offset = 0;
            ntt_columns(tables_, stream_, d_combined, L, d_ccoef, L, d_tmp, 1, log_L, true, &sc);
            check_all_zero(stream_, d_ccoef + 7 * n, n, d_flag_);
            d2h(&flag, d_flag_, sizeof(flag));
        }
        if (flag & 4) throw ProveFailure{EZK_ERR_INVALID_ARGUMENT, "the operation list contains an unknown opcode"};
        if (flag & 2)
            throw ProveFailure{EZK_ERR_INVALID_ARGUMENT, "the trace contains non-canonical field elements (values >= the modulus)"};
        if (flag)
            throw ProveFailure{EZK_ERR_CONSTRAINT_DEGREE,
                               "constraint composition polynomial has degree >= 7n: the trace does not satisfy the AIR"};
        if (sharded) {
            const uint64_t slice = n >> glog;
            EZK_CUDA(cudaEventRecord(share_ev_[0], stream_));
            EZK_CUDA(cudaStreamWaitEvent(aux_stream_, share_ev_[0], 0));
            for (uint32_t j = 0; j < kCompCols; j++) {
                comm_.all_gather(d_rloc + (size_t)j * slice, d_ccoef + (size_t)j * n, slice * 16, aux_stream_);
                count_launch();
                EZK_CUDA(cudaEventRecord(share_ev_[1 + j], aux_stream_));
            }
            const uint32_t cuts[4] = {0, 2, 4, kCompCols};  // extension launches of 2, 2 and 3 columns
            for (int b = 0; b < 3; b++) {
                EZK_CUDA(cudaStreamWaitEvent(stream_, share_ev_[cuts[b + 1]], 0));
                lde_columns(tables_, stream_, d_ccoef + (size_t)cuts[b] * n, n, d_clde + (size_t)cuts[b] * L_local, L_local, d_tmp,
                            cuts[b + 1] - cuts[b], log_n, cs);
            }
        } else {
            for (uint32_t c0 = 0; c0 < kCompCols; c0 += comp_group)
                lde_columns(tables_, stream_, d_ccoef + (size_t)c0 * n, n, d_clde + (size_t)c0 * L, L, d_tmp,
                            std::min(comp_group, kCompCols - c0), log_n, cs);
        }
        last_.comp_root = commit(d_clde, kCompCols, L, d_cnodes, comp_tree);
        commitments.insert(commitments.end(), last_.comp_root.begin(), last_.comp_root.end());
        coin.reseed(last_.comp_root);
    }
    mark();

    // ---- (4) out-of-domain frame + DEEP composition ----
    const Fp z = coin.draw(), zg = z * g;
    std::vector<Fp> ood_trace(2 * kWidth), ood_comp(kCompCols);  // ood_trace interleaved [cur_c, next_c]
    {
        uint64_t y[2][2];
        put(y[0], z * o_inv), put(y[1], zg * o_inv);
        uint4* d_ood = d_small;  // 56 + 7 elements
        {
            // 1 / ((x - z)(x - zg)) needs z only: auxiliary stream, overlapped with the OOD evaluation, its trip to
            // the host and the coefficient-space combination
            EZK_CUDA(cudaEventRecord(aux_ev_[2], stream_));  // d_invden is free once the constraint kernel is done
            EZK_CUDA(cudaStreamWaitEvent(aux_stream_, aux_ev_[2], 0));
            uint64_t a[2], b[2];
            put(a, z), put(b, zg);
            domain_pair_inverse(aux_stream_, tables_.root_fwd, tables_.root_inv, log_L, a, b, d_invden, sh);
            EZK_CUDA(cudaEventRecord(aux_ev_[3], aux_stream_));
        }
        // powers of the two points once (d_pq is free until the DEEP combination), then 63 dot products
        power_table(stream_, log_n, y, 2, d_pq);
        if (sharded) {
            // rank r evaluates the trace columns c = r mod G and the composition columns j = r mod G; the blocks of
            // 2 * rounds + ceil(7 / G) values are all-gathered
            const uint32_t cr = (kCompCols + G - 1) / G, block = 2 * rounds + cr;
            const uint32_t mine_t = me < kWidth ? (kWidth - me + G - 1) / G : 0, mine_c = me < kCompCols ? (kCompCols - me + G - 1) / G : 0;
            if (mine_t) eval_polys(stream_, d_tcoef + (size_t)me * n, (uint64_t)G * n, mine_t, log_n, d_pq, 2, d_scratch, d_ood);
            if (mine_c)
                eval_polys(stream_, d_ccoef + (size_t)me * n, (uint64_t)G * n, mine_c, log_n, d_pq, 1,
                           d_scratch + (size_t)2 * kWidth * eval_blocks, d_ood + 2 * rounds);
            uint4* d_ood_all = d_small + 128;
            comm_.all_gather(d_ood, d_ood_all, (size_t)block * 16, stream_);
            count_launch();
            std::vector<Fp> host((size_t)block * G);
            d2h(host.data(), d_ood_all, host.size() * 16);
            for (uint32_t c = 0; c < kWidth; c++)
                for (uint32_t k = 0; k < 2; k++) ood_trace[2 * c + k] = host[(size_t)(c % G) * block + 2 * (c / G) + k];
            for (uint32_t j = 0; j < kCompCols; j++) ood_comp[j] = host[(size_t)(j % G) * block + 2 * rounds + j / G];
        } else {
            eval_polys(stream_, d_tcoef, n, kWidth, log_n, d_pq, 2, d_scratch, d_ood);
            eval_polys(stream_, d_ccoef, n, kCompCols, log_n, d_pq, 1, d_scratch + (size_t)2 * kWidth * eval_blocks, d_ood + 2 * kWidth);
            std::vector<Fp> host(2 * kWidth + kCompCols);
            d2h(host.data(), d_ood, host.size() * 16);
            std::copy(host.begin(), host.begin() + 2 * kWidth, ood_trace.begin());
            std::copy(host.begin() + 2 * kWidth, host.end(), ood_comp.begin());
        }
    }
    const WireCompat wc = wire_compat();
    // the frame as it is hashed and serialized: interleaved, or all current states followed by all next states
    std::vector<Fp> ood_wire = ood_trace;
    if (!wc.ood_interleaved)
        for (uint32_t c = 0; c < kWidth; c++) ood_wire[c] = ood_trace[2 * c], ood_wire[kWidth + c] = ood_trace[2 * c + 1];
    coin.reseed(hash_elements(ood_wire.data(), ood_wire.size()));
    coin.reseed(hash_elements(ood_comp.data(), ood_comp.size()));
    last_.ood_trace = ood_trace, last_.ood_comp = ood_comp;
    uint4* d_deep_local = d_deep;  // sharded: this rank's DEEP evaluations in packed row order (L / G of them)
    {
        std::vector<Fp> dc(kWidth + kCompCols);
        for (auto& v : dc) v = coin.draw();
        Fp s1, s2;
        for (uint32_t c = 0; c < kWidth; c++) {
            s1 = s1 + dc[c] * ood_trace[2 * c];
            s2 = s2 + dc[c] * ood_trace[2 * c + 1];
        }
        for (uint32_t j = 0; j < kCompCols; j++) s1 = s1 + dc[kWidth + j] * ood_comp[j];
        uint4* d_dc = d_small + 512;
        h2d(d_dc, dc.data(), dc.size() * 16);
        DeepScalars ds;
        put(ds.z, z), put(ds.zg, zg), put(ds.s1, s1), put(ds.s2, s2);
        if (sharded) {
            // the pointwise formula on the rows this rank owns: no communication
            EZK_CUDA(cudaStreamWaitEvent(stream_, aux_ev_[3], 0));
            deep_from_rows(stream_, tables_.root_fwd, d_tlde, L_local, d_clde, L_local, log_L, d_dc, d_invden, ds, d_deep_local, sh);
        } else {
            deep_combine_coeffs(stream_, d_tcoef, n, d_ccoef, n, log_n, d_dc, d_pq);
            lde_columns(tables_, stream_, d_pq, n, d_pqlde, L, d_tmp, 2, log_n, cs);
            EZK_CUDA(cudaStreamWaitEvent(stream_, aux_ev_[3], 0));
            deep_pointwise(stream_, tables_.root_fwd, d_pqlde, log_L, d_invden, ds, d_deep, sh);
        }
    }
    mark();

    // ---- (5) FRI ----
    struct Layer {
        uint4* evals;       // whole layer in natural order, or (packed) the positions p = me mod G in ascending order
        bool packed;
        uint64_t size;      // of the whole layer
        ShardTree tree;
    };
    std::vector<Layer> layers;
    layers.reserve(nlayers);
    std::vector<Fp> remainder;
    {
        uint4* cur = d_deep_local;
        bool cur_packed = sharded;
        uint64_t s = L;
        // the positions every rank holds, put back into natural order on every rank (the tail of FRI is replicated)
        auto gather_natural = [&](uint4* packed, uint64_t size) {
            uint4* natural = alloc(size);
            comm_.all_gather(packed, d_allg, (size >> glog) * 16, stream_);
            count_launch();
            unpack_rows(stream_, d_allg, size >> glog, glog, 1, natural);
            return natural;
        };
        FriFoldConsts fc{};
        const Fp zinv = inverse(root_of_unity(3));
        Fp zp(1);
        for (int k = 0; k < 4; k++) {
            put(fc.zinv[k], zp);
            zp = zp * zinv;
        }
        put(fc.inv8, inverse(Fp::from_u64(8)));
        for (size_t k = 0; k < nlayers; k++) {
            const uint64_t m = s / 8;
            const bool split = cur_packed && m >= shard_fri_min_rows && m >= (uint64_t)G * G;
            if (cur_packed && !split) cur = gather_natural(cur, s), cur_packed = false;
            layers.push_back(Layer{cur, split, s, ShardTree{}});
            Layer& layer = layers.back();
            Hash32 root;
            if (split) {
                // a row is 8 positions at stride m, all of them = i mod G: in packed order again 8 values at stride m / G
                root = commit(cur, 8, m, nullptr, layer.tree);
            } else {
                uint4* nodes = alloc(4 * m);
                const bool was = sharded;
                sharded = false;  // whole-layer tree on every rank
                root = commit(cur, 8, m, nodes, layer.tree);
                sharded = was;
            }
            commitments.insert(commitments.end(), root.begin(), root.end());
            coin.reseed(root);
            last_.fri_roots.push_back(root);
            const Fp alpha = coin.draw();
            put(fc.alpha_oinv, alpha * o_inv);
            uint4* next = alloc(split ? (m >> glog) : m);
            fri_fold(stream_, tables_.root_inv, cur, ilog2_u64(s), fc, next, split ? sh : RowShard());
            cur = next, s = m;
        }
        if (cur_packed) cur = gather_natural(cur, s), cur_packed = false;
        // remainder: interpolate the last layer over the coset; only the first s/8 coefficients may be non-zero
        uint4* d_rem = alloc(s);
        {
            // interpolation over the coset 3 * <w_s>: inverse transform, then coefficient k times 3^-k / s (the batched
            // NTT kernels; the layer is at most 4096 points)
            NttScale sc{};
            put(sc.cvec[0], inverse(Fp::from_u64(s)));
            sc.chunk_shift = 63, sc.use_offset = 2;
            ntt_columns(tables_, stream_, cur, s, d_rem, s, d_tmp, 1, ilog2_u64(s), true, &sc);
        }
        std::vector<Fp> all(s);
        d2h(all.data(), d_rem, s * 16);
        for (uint64_t i = s / 8; i < s; i++)
            if (!all[i].is_zero())
                throw ProveFailure{EZK_ERR_DEEP_DEGREE, "FRI remainder has degree >= domain/8: DEEP composition degree too high"};
        remainder.assign(all.begin(), all.begin() + s / 8);
        if (!wc.remainder_low_to_high) std::reverse(remainder.begin(), remainder.end());
        Hash32 commitment = hash_elements(remainder.data(), remainder.size());
        commitments.insert(commitments.end(), commitment.begin(), commitment.end());
        coin.reseed(commitment);
        last_.fri_roots.push_back(commitment);
        last_.remainder = remainder;
    }
    mark();

    // ---- (6) grinding + query positions ----
    uint64_t nonce = wc.first_nonce;
    while (coin.leading_zeros(nonce) < opt.grinding) nonce++;
    std::vector<uint64_t> positions = coin.draw_integers(opt.num_queries, L, nonce);
    std::sort(positions.begin(), positions.end());
    positions.erase(std::unique(positions.begin(), positions.end()), positions.end());
    last_.positions = positions;

    // ---- (7) openings + serialization ----
    ProofWriter w;
    // TraceInfo: main width, aux width, aux random elements, log2(length), metadata length
    w.u8((uint8_t)kWidth), w.u8(0);
    if (wc.trace_info_aux_rands_byte) w.u8(0);
    w.u8((uint8_t)log_n), w.u16(0);
    w.u8(16);
    w.element(Fp(Fp::modulus()));
    w.u8((uint8_t)opt.num_queries), w.u8((uint8_t)opt.blowup), w.u8((uint8_t)opt.grinding), w.u8((uint8_t)opt.field_ext);
    w.u8((uint8_t)opt.fri_fold), w.u8((uint8_t)opt.fri_rem_max_deg);
    w.u8((uint8_t)positions.size());
    w.u16((uint16_t)commitments.size());
    w.bytes(commitments.data(), commitments.size());

    // All openings (trace rows, constraint rows, one per FRI layer) are planned on the host first, fetched with ONE
    // index upload, a burst of gather kernels and ONE read-back, and serialized afterwards.  Multi-GPU: every rank
    // plans the same openings, gathers the rows and subtree nodes it owns (any index elsewhere), the gather buffers
    // are all-gathered and each item is taken from its owner's copy; the top log2(G) tree levels live on the host.
    uint64_t* d_idx = reinterpret_cast<uint64_t*>(alloc(8192));  // up to 16384 u64 indices
    uint4* d_gather = alloc(32768);
    struct Opening {
        const uint4* table;
        uint64_t pitch;
        uint32_t width;
        const ShardTree* tree;
        std::vector<uint64_t> pos;
        std::vector<uint32_t> row_owner;                   // rank whose gather buffer holds row q
        std::vector<std::vector<uint64_t>> idx_lists;      // global node indices, serialization order
        std::vector<int32_t> dig_owner;                    // per digest: owning rank, or -1 = host-side top levels
        std::vector<uint64_t> dig_top;                     // per digest: index into tree->top when dig_owner < 0
        size_t idx_off, ndig, out_off;  // offsets into the index array / the gather buffer (16-byte units)
    };
    std::vector<Opening> openings;
    openings.reserve(2 + nlayers);
    std::vector<uint64_t> flat;
    size_t out_units = 0;
    // rows_mode 0: whole table on this rank; 1: row p is owned by rank p mod G and sits at index p / G there (packed)
    auto plan_opening = [&](const uint4* table, uint64_t pitch, uint32_t width, const ShardTree* tree, int rows_mode,
                            const std::vector<uint64_t>& pos) {
        Opening op{table, pitch, width, tree, pos, {}, batch_proof_node_indices(tree->leaves, pos), {}, {}, flat.size(), 0, out_units};
        for (uint64_t p : pos) {
            const uint32_t owner = rows_mode == 0 ? me : (uint32_t)(p & (G - 1));
            op.row_owner.push_back(owner);
            flat.push_back(owner != me ? 0 : rows_mode == 1 ? p >> glog : p);
        }
        for (auto& v : op.idx_lists)
            for (uint64_t k : v) {
                if (!tree->split) {
                    op.dig_owner.push_back((int32_t)me), op.dig_top.push_back(0), flat.push_back(k);
                } else {
                    const NodeHome home = shard_node_home(G, glog, k);  // subtree of a rank, or the host-side top levels
                    op.dig_owner.push_back(home.owner), op.dig_top.push_back(home.owner < 0 ? home.index : 0);
                    flat.push_back(home.owner == (int)me ? home.index : 1);
                }
            }
        op.ndig = op.dig_owner.size();
        out_units += pos.size() * width + 2 * op.ndig;
        openings.push_back(std::move(op));
    };
    plan_opening(d_tlde, L_local, kWidth, &trace_tree, sharded ? 1 : 0, positions);
    plan_opening(d_clde, L_local, kCompCols, &comp_tree, sharded ? 1 : 0, positions);
    {
        std::vector<uint64_t> pos = positions;
        for (auto& layer : layers) {
            pos = fold_positions(pos, layer.size, 8);
            const uint64_t m = layer.size / 8;
            plan_opening(layer.evals, layer.packed ? m >> glog : m, 8, &layer.tree, layer.packed ? 1 : 0, pos);
        }
    }
    const bool exchange = sharded && G > 1;
    if (flat.size() > 16384 || out_units > 32768 || out_units * 16 * (exchange ? G : 1) > pinned_bytes_)
        throw ProveFailure{EZK_ERR_INTERNAL, "query staging too small"};
    h2d(d_idx, flat.data(), flat.size() * 8);
    for (auto& op : openings) {
        const uint32_t nq = (uint32_t)op.pos.size();
        gather_rows(stream_, op.table, op.pitch, op.width, d_idx + op.idx_off, nq, d_gather + op.out_off);
        if (op.ndig)
            gather_digests(stream_, op.tree->nodes, d_idx + op.idx_off + nq, (uint32_t)op.ndig, d_gather + op.out_off + (size_t)nq * op.width);
    }
    std::vector<uint8_t> fetched(out_units * 16), all;
    if (exchange) {
        comm_.all_gather(d_gather, d_allg, out_units * 16, stream_);
        count_launch();
        all.resize(out_units * 16 * G);
        d2h(all.data(), d_allg, all.size());
    } else {
        d2h(fetched.data(), d_gather, fetched.size());
    }
    auto item = [&](uint32_t owner, size_t unit) -> const uint8_t* {  // `unit` = offset in 16-byte words in a gather buffer
        return exchange ? all.data() + ((size_t)owner * out_units + unit) * 16 : fetched.data() + unit * 16;
    };
    auto write_opening = [&](const Opening& op) {
        const size_t nq = op.pos.size();
        w.u32((uint32_t)(nq * op.width * 16));
        for (size_t q = 0; q < nq; q++) w.bytes(item(op.row_owner[q], op.out_off + q * op.width), (size_t)op.width * 16);
        std::vector<uint8_t> paths;
        paths.push_back((uint8_t)op.idx_lists.size());
        size_t d = 0;
        for (auto& v : op.idx_lists) {
            paths.push_back((uint8_t)v.size());
            for (size_t k = 0; k < v.size(); k++, d++) {
                const uint8_t* src = op.dig_owner[d] < 0 ? op.tree->top[op.dig_top[d]].data()
                                                         : item((uint32_t)op.dig_owner[d], op.out_off + nq * op.width + 2 * d);
                paths.insert(paths.end(), src, src + 32);
            }
        }
        w.u32((uint32_t)paths.size());
        w.bytes(paths.data(), paths.size());
    };
    write_opening(openings[0]);
    write_opening(openings[1]);
    // OodFrame
    w.u16((uint16_t)(1 + ood_wire.size() * 16));
    w.u8(2);
    for (Fp v : ood_wire) w.element(v);
    w.u16(1), w.u8(0);
    w.u16((uint16_t)(ood_comp.size() * 16));
    for (Fp v : ood_comp) w.element(v);
    // FriProof
    w.u8((uint8_t)nlayers);
    for (size_t k = 2; k < openings.size(); k++) write_opening(openings[k]);
    w.u16((uint16_t)(remainder.size() * 16));
    for (Fp v : remainder) w.element(v);
    w.u8(1);
    w.u64(nonce);
    w.u8(0);
    mark();
    sync();
    for (int i = 0; i + 1 < evi && i < 8; i++) EZK_CUDA(cudaEventElapsedTime(&stage_ms_[i], ev_[i], ev_[i + 1]));
    return std::move(w.data());
}

std::vector<uint8_t> GpuProver::artifact(int which) {
    EZK_CUDA(cudaSetDevice(device_));
    auto from_device = [&](const uint4* p, size_t elems) {
        std::vector<uint8_t> out(elems * 16);
        if (!p) throw ProveFailure{EZK_ERR_INVALID_ARGUMENT, "no proof has been generated yet"};
        sync();
        EZK_CUDA(cudaMemcpy(out.data(), p, out.size(), cudaMemcpyDeviceToHost));
        return out;
    };
    auto from_elems = [](const std::vector<Fp>& v) {
        return std::vector<uint8_t>((const uint8_t*)v.data(), (const uint8_t*)(v.data() + v.size()));
    };
    switch (which) {
        case EZK_ART_TRACE_ROOT: return std::vector<uint8_t>(last_.trace_root.begin(), last_.trace_root.end());
        case EZK_ART_CONSTRAINT_ROOT: return std::vector<uint8_t>(last_.comp_root.begin(), last_.comp_root.end());
        case EZK_ART_COMBINED: return from_device(last_.combined_copy, last_.L);
        case EZK_ART_OOD_TRACE: return from_elems(last_.ood_trace);
        case EZK_ART_OOD_CONSTRAINTS: return from_elems(last_.ood_comp);
        case EZK_ART_DEEP_EVALS: return from_device(last_.deep, last_.L);
        case EZK_ART_FRI_ROOTS: {
            std::vector<uint8_t> out;
            for (auto& r : last_.fri_roots) out.insert(out.end(), r.begin(), r.end());
            return out;
        }
        case EZK_ART_REMAINDER: return from_elems(last_.remainder);
        case EZK_ART_POSITIONS:
            return std::vector<uint8_t>((const uint8_t*)last_.positions.data(),
                                        (const uint8_t*)(last_.positions.data() + last_.positions.size()));
        case EZK_ART_TRACE_LDE: return from_device(last_.tlde, kWidth * last_.L);
        case EZK_ART_CONSTRAINT_LDE: return from_device(last_.clde, kCompCols * last_.L);
        case EZK_ART_TRACE_POLYS: return from_device(last_.tcoef, kWidth * last_.n);
    }
    throw ProveFailure{EZK_ERR_INVALID_ARGUMENT, "unknown artifact id"};
}

// ---------------------------------------------------------------------------------------------------------
// stage-level helpers

void GpuProver::stage_lde(const void* columns, uint32_t width, uint64_t n, void* lde_out) {
    if (n < 8 || (n & (n - 1)) || width == 0) throw ProveFailure{EZK_ERR_INVALID_ARGUMENT, "bad LDE shape"};
    EZK_CUDA(cudaSetDevice(device_));
    const uint64_t L = 8 * n;
    const uint32_t log_n = ilog2_u64(n);
    reserve((size_t)width * (2 * n + 2 * L) + 1024);
    reset_arena();
    last_ = Last{};
    uint4 *d_in = alloc(width * n), *d_coef = alloc(width * n), *d_lde = alloc(width * L), *d_tmp = alloc(width * L);
    EZK_CUDA(cudaMemcpyAsync(d_in, columns, width * n * 16, cudaMemcpyHostToDevice, stream_));
    NttScale sc{};
    put(sc.cvec[0], inverse(Fp::from_u64(n)));
    sc.chunk_shift = 63, sc.use_offset = 1;
    ntt_columns(tables_, stream_, d_in, n, d_coef, n, d_tmp, width, log_n, true, &sc);
    lde_columns(tables_, stream_, d_coef, n, d_lde, L, d_tmp, width, log_n);
    EZK_CUDA(cudaMemcpyAsync(lde_out, d_lde, width * L * 16, cudaMemcpyDeviceToHost, stream_));
    sync();
}

void GpuProver::stage_ntt(const void* columns, uint32_t width, uint64_t n, bool inv, void* out) {
    if (n < 2 || (n & (n - 1)) || width == 0) throw ProveFailure{EZK_ERR_INVALID_ARGUMENT, "bad NTT shape"};
    EZK_CUDA(cudaSetDevice(device_));
    const uint32_t log_n = ilog2_u64(n);
    reserve((size_t)width * 3 * n + 1024);
    reset_arena();
    last_ = Last{};
    uint4 *d_in = alloc(width * n), *d_out = alloc(width * n), *d_tmp = alloc(width * n);
    EZK_CUDA(cudaMemcpyAsync(d_in, columns, width * n * 16, cudaMemcpyHostToDevice, stream_));
    NttScale sc{};
    put(sc.cvec[0], inv ? inverse(Fp::from_u64(n)) : Fp(1));
    sc.chunk_shift = 63, sc.use_offset = 0;
    ntt_columns(tables_, stream_, d_in, n, d_out, n, d_tmp, width, log_n, inv, inv ? &sc : nullptr);
    EZK_CUDA(cudaMemcpyAsync(out, d_out, width * n * 16, cudaMemcpyDeviceToHost, stream_));
    sync();
}

void GpuProver::stage_merkle(const void* table, uint32_t width, uint64_t rows, void* nodes_out) {
    if (rows < 2 || (rows & (rows - 1)) || width == 0) throw ProveFailure{EZK_ERR_INVALID_ARGUMENT, "bad Merkle shape"};
    EZK_CUDA(cudaSetDevice(device_));
    reserve((size_t)width * rows + 4 * rows + 1024);
    reset_arena();
    last_ = Last{};
    uint4 *d_table = alloc(width * rows), *d_nodes = alloc(4 * rows);
    EZK_CUDA(cudaMemcpyAsync(d_table, table, width * rows * 16, cudaMemcpyHostToDevice, stream_));
    EZK_CUDA(cudaMemsetAsync(d_nodes, 0, 64, stream_));  // node 0 is unused
    merkle_hash_rows(stream_, d_table, rows, width, rows, d_nodes);
    merkle_build(stream_, d_nodes, rows);
    EZK_CUDA(cudaMemcpyAsync(nodes_out, d_nodes, 2 * rows * 32, cudaMemcpyDeviceToHost, stream_));
    sync();
}

void GpuProver::stage_fri_fold(const void* evals, uint64_t s, Fp alpha, void* next_out) {
    if (s < 16 || (s & (s - 1))) throw ProveFailure{EZK_ERR_INVALID_ARGUMENT, "bad FRI layer size"};
    EZK_CUDA(cudaSetDevice(device_));
    reserve(s + s / 8 + 1024);
    reset_arena();
    last_ = Last{};
    uint4 *d_e = alloc(s), *d_n = alloc(s / 8);
    EZK_CUDA(cudaMemcpyAsync(d_e, evals, s * 16, cudaMemcpyHostToDevice, stream_));
    FriFoldConsts fc{};
    const Fp zinv = inverse(root_of_unity(3));
    Fp zp(1);
    for (int k = 0; k < 4; k++) {
        put(fc.zinv[k], zp);
        zp = zp * zinv;
    }
    put(fc.inv8, inverse(Fp::from_u64(8)));
    put(fc.alpha_oinv, alpha * inverse(Fp::from_u64(kDomainOffset)));
    fri_fold(stream_, tables_.root_inv, d_e, ilog2_u64(s), fc, d_n);
    EZK_CUDA(cudaMemcpyAsync(next_out, d_n, (s / 8) * 16, cudaMemcpyDeviceToHost, stream_));
    sync();
}

void GpuProver::stage_eval_frames_sum(const void* cur, const void* nxt, const void* periodic, uint32_t nframes, uint32_t delta,
                                      const Fp tcoef[20], void* out1) {
    EZK_CUDA(cudaSetDevice(device_));
    reserve((size_t)nframes * (28 * 2 + 9 + 1) + 1024);
    reset_arena();
    last_ = Last{};
    uint4 *d_c = alloc(nframes * 28), *d_n = alloc(nframes * 28), *d_p = alloc(nframes * 9), *d_o = alloc(nframes);
    EZK_CUDA(cudaMemcpyAsync(d_c, cur, (size_t)nframes * 28 * 16, cudaMemcpyHostToDevice, stream_));
    EZK_CUDA(cudaMemcpyAsync(d_n, nxt, (size_t)nframes * 28 * 16, cudaMemcpyHostToDevice, stream_));
    EZK_CUDA(cudaMemcpyAsync(d_p, periodic, (size_t)nframes * 9 * 16, cudaMemcpyHostToDevice, stream_));
    ConstraintParams hp{};
    hp.delta = delta;
    for (uint32_t j = 0; j < kTransitions; j++) put(hp.tcoef[j], tcoef[j]), put_pre(hp.tcoef_pre[j], tcoef[j]);
    for (int i = 0; i < 16; i++) put(hp.inv_mds[i], rescue_const(rescue_inv_mds()[i])), put_pre(hp.inv_mds_pre[i], rescue_const(rescue_inv_mds()[i]));
    EZK_CUDA(cudaMemcpyAsync(d_params_, &hp, sizeof(hp), cudaMemcpyHostToDevice, stream_));
    evaluate_frames_sum(stream_, d_c, d_n, d_p, nframes, d_params_, d_o);
    EZK_CUDA(cudaMemcpyAsync(out1, d_o, (size_t)nframes * 16, cudaMemcpyDeviceToHost, stream_));
    sync();
}

void GpuProver::stage_eval_frames(const void* cur, const void* nxt, const void* periodic, uint32_t nframes, uint32_t delta,
                                  void* out20) {
    EZK_CUDA(cudaSetDevice(device_));
    reserve((size_t)nframes * (28 * 2 + 9 + 20) + 1024);
    reset_arena();
    last_ = Last{};
    uint4 *d_c = alloc(nframes * 28), *d_n = alloc(nframes * 28), *d_p = alloc(nframes * 9), *d_o = alloc(nframes * 20);
    EZK_CUDA(cudaMemcpyAsync(d_c, cur, (size_t)nframes * 28 * 16, cudaMemcpyHostToDevice, stream_));
    EZK_CUDA(cudaMemcpyAsync(d_n, nxt, (size_t)nframes * 28 * 16, cudaMemcpyHostToDevice, stream_));
    EZK_CUDA(cudaMemcpyAsync(d_p, periodic, (size_t)nframes * 9 * 16, cudaMemcpyHostToDevice, stream_));
    ConstraintParams hp{};
    hp.delta = delta;
    for (int i = 0; i < 16; i++) put(hp.inv_mds[i], rescue_const(rescue_inv_mds()[i])), put_pre(hp.inv_mds_pre[i], rescue_const(rescue_inv_mds()[i]));
    EZK_CUDA(cudaMemcpyAsync(d_params_, &hp, sizeof(hp), cudaMemcpyHostToDevice, stream_));
    evaluate_frames(stream_, d_c, d_n, d_p, nframes, d_params_, d_o);
    EZK_CUDA(cudaMemcpyAsync(out20, d_o, (size_t)nframes * 20 * 16, cudaMemcpyDeviceToHost, stream_));
    sync();
}

void GpuProver::bench_lde_merkle(uint32_t width, uint64_t n, int iters, float* lde_ms, float* merkle_ms) {
    if (n < 8 || (n & (n - 1)) || width == 0 || iters < 1) throw ProveFailure{EZK_ERR_INVALID_ARGUMENT, "bad shape"};
    EZK_CUDA(cudaSetDevice(device_));
    const uint64_t L = 8 * n;
    const uint32_t log_n = ilog2_u64(n);
    reserve((size_t)width * (2 * n + 2 * L) + 4 * L + 1024);
    reset_arena();
    last_ = Last{};
    uint4 *d_in = alloc(width * n), *d_coef = alloc(width * n), *d_lde = alloc(width * L), *d_tmp = alloc(width * L);
    uint4* d_nodes = alloc(4 * L);
    // synthetic column data: any byte pattern below the modulus will do (top word cleared)
    EZK_CUDA(cudaMemsetAsync(d_in, 0x5A, width * n * 16, stream_));
    NttScale sc{};
    put(sc.cvec[0], inverse(Fp::from_u64(n)));
    sc.chunk_shift = 63, sc.use_offset = 1;
    float t_lde = 0, t_mk = 0;
    for (int it = -1; it < iters; it++) {  // iteration -1 is a warm-up
        EZK_CUDA(cudaEventRecord(ev_[0], stream_));
        ntt_columns(tables_, stream_, d_in, n, d_coef, n, d_tmp, width, log_n, true, &sc);
        lde_columns(tables_, stream_, d_coef, n, d_lde, L, d_tmp, width, log_n);
        EZK_CUDA(cudaEventRecord(ev_[1], stream_));
        merkle_hash_rows(stream_, d_lde, L, width, L, d_nodes);
        merkle_build(stream_, d_nodes, L);
        EZK_CUDA(cudaEventRecord(ev_[2], stream_));
        sync();
        float a, b;
        EZK_CUDA(cudaEventElapsedTime(&a, ev_[0], ev_[1]));
        EZK_CUDA(cudaEventElapsedTime(&b, ev_[1], ev_[2]));
        if (it >= 0) t_lde += a, t_mk += b;
    }
    *lde_ms = t_lde / iters, *merkle_ms = t_mk / iters;
}

void GpuProver::bench_fri(uint64_t n, int iters, float* fri_ms) {
    if (n < 64 || (n & (n - 1)) || iters < 1) throw ProveFailure{EZK_ERR_INVALID_ARGUMENT, "bad shape"};
    EZK_CUDA(cudaSetDevice(device_));
    const uint64_t L = 8 * n;
    ProofOptions opt;
    const size_t nlayers = num_fri_layers(L, opt);
    reserve(L + L + 1024);
    reset_arena();
    last_ = Last{};
    uint4* d_e = alloc(L);
    EZK_CUDA(cudaMemsetAsync(d_e, 0x5A, L * 16, stream_));
    FriFoldConsts fc{};
    const Fp zinv = inverse(root_of_unity(3));
    Fp zp(1);
    for (int k = 0; k < 4; k++) {
        put(fc.zinv[k], zp);
        zp = zp * zinv;
    }
    put(fc.inv8, inverse(Fp::from_u64(8)));
    put(fc.alpha_oinv, Fp::from_u64(0x123456789ULL));
    const size_t mark0 = arena_.used;
    float total = 0;
    for (int it = -1; it < iters; it++) {
        arena_.used = mark0;
        uint4* cur = d_e;
        uint64_t s = L;
        EZK_CUDA(cudaEventRecord(ev_[0], stream_));
        for (size_t k = 0; k < nlayers; k++) {
            const uint64_t m = s / 8;
            uint4* nodes = alloc(4 * m);
            merkle_hash_rows(stream_, cur, m, 8, m, nodes);
            merkle_build(stream_, nodes, m);
            uint4* next = alloc(m);
            fri_fold(stream_, tables_.root_inv, cur, ilog2_u64(s), fc, next);
            cur = next, s = m;
        }
        EZK_CUDA(cudaEventRecord(ev_[1], stream_));
        sync();
        float a;
        EZK_CUDA(cudaEventElapsedTime(&a, ev_[0], ev_[1]));
        if (it >= 0) total += a;
    }
    *fri_ms = total / iters;
}

}  // namespace ezk
