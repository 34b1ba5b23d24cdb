// extern "C" boundary (include/ezkvm_prover.h).  Plain pointers and sizes only; every C++ exception is
// translated into an ezk_status code + thread-local message.
#include "../../include/ezkvm_prover.h"
#include "common.h"
#include "host/vm.h"
#include "host/launch_groups.h"
#include "dist/shard_layout.h"
#include "prover.h"
#include "trace/expand.cuh"
#include <algorithm>
#include <cstring>
#include <memory>
#include <mutex>
#include <vector>

using namespace ezk;

struct ezk_prover {
    std::unique_ptr<GpuProver> impl;
    std::mutex mu;
};
struct ezk_group {
    std::unique_ptr<LocalGroup> impl;
};
struct ezk_program {
    Program prog;
};
struct ezk_execution {
    ExecutionTrace trace;
};

namespace {
thread_local std::string g_error;

template <class F>
int guarded(F&& f) {
    try {
        f();
        return EZK_OK;
    } catch (const ProveFailure& e) {
        g_error = e.message;
        return e.code;
    } catch (const VmError& e) {
        g_error = e.message;
        return EZK_ERR_VM;
    } catch (const CudaError& e) {
        g_error = e.what();
        return EZK_ERR_CUDA;
    } catch (const std::bad_alloc&) {
        g_error = "out of host memory";
        return EZK_ERR_INTERNAL;
    } catch (const std::exception& e) {
        g_error = e.what();
        return EZK_ERR_INTERNAL;
    }
}

ProofOptions to_options(const ezk_options* o) {
    ProofOptions p;
    if (o) {
        p.num_queries = o->num_queries, p.blowup = o->blowup_factor, p.grinding = o->grinding_factor;
        p.field_ext = o->field_extension, p.fri_fold = o->fri_folding_factor, p.fri_rem_max_deg = o->fri_remainder_max_degree;
    }
    return p;
}

PublicInputs to_public(const ezk_public_inputs* pi) {
    if (!pi) throw ProveFailure{EZK_ERR_INVALID_ARGUMENT, "public inputs are required"};
    PublicInputs p;
    for (int i = 0; i < 2; i++) p.elements[i] = fp_load(pi->program_hash[i]);
    for (int i = 0; i < 16; i++) p.elements[2 + i] = fp_load(pi->stack_outputs[i]);
    for (int i = 0; i < 18; i++)
        if (p.elements[i].v >= Fp::modulus()) throw ProveFailure{EZK_ERR_INVALID_ARGUMENT, "public input is not a canonical field element"};
    p.lwe_k = pi->lwe_k, p.lwe_delta = pi->lwe_delta;
    return p;
}

void export_proof(std::vector<uint8_t>& bytes, uint8_t** proof, size_t* proof_len) {
    uint8_t* out = static_cast<uint8_t*>(malloc(bytes.size()));
    if (!out) throw std::bad_alloc();
    memcpy(out, bytes.data(), bytes.size());
    *proof = out;
    *proof_len = bytes.size();
}

ezk_prover* g_default = nullptr;
std::mutex g_default_mu;
}  // namespace

extern "C" {

const char* ezk_last_error(void) { return g_error.c_str(); }
const char* ezk_version(void) { return "encrypt-zkvm-b200 0.1 (sm_100a)"; }
int ezk_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}
uint64_t ezk_kernel_launch_count(void) { return launch_count(); }
void ezk_free(void* p) { free(p); }
int ezk_selftest_copy_pool(uint32_t threads, size_t bytes) {
    return guarded([&] {
        if (threads == 0 || threads > 64 || bytes > ((size_t)1 << 30)) throw ProveFailure{EZK_ERR_INVALID_ARGUMENT, "bad self-test size"};
        std::vector<uint8_t> src(bytes + 64), dst(bytes + 64);
        uint64_t x = 0x9E3779B97F4A7C15ULL ^ bytes;
        for (auto& b : src) b = (uint8_t)((x = x * 6364136223846793005ULL + 1442695040888963407ULL) >> 56);
        CopyPool pool(threads);
        const size_t lens[] = {bytes, bytes / 2 + 1, bytes > 4096 ? bytes - 4095 : bytes, 262145 < bytes ? 262145 : bytes, 0};
        for (size_t len : lens)
            for (size_t shift : {(size_t)0, (size_t)1, (size_t)63}) {
                std::fill(dst.begin(), dst.end(), (uint8_t)0xEE);
                pool.copy(dst.data() + shift, src.data() + shift, len);
                if (memcmp(dst.data() + shift, src.data() + shift, len) != 0 || dst[shift + len] != 0xEE || (shift && dst[shift - 1] != 0xEE))
                    throw ProveFailure{EZK_ERR_INTERNAL, "threaded copy differs from memcpy"};
            }
    });
}
// Discrete-event replay of host/launch_groups.h: column c arrives at (c + 1) * upload_us, a launch of k columns takes
// k * compute_us and starts when its last column is there and the launch before it has ended; the decision is taken
// whenever a column arrives (the staged upload) and, once all have arrived, whenever a launch ends.
int ezk_selftest_launch_groups(uint32_t columns, uint32_t cap, uint32_t upload_us, uint32_t compute_us, uint32_t* sizes_out,
                               uint32_t* groups_out, uint64_t* idle_us_out) {
    return guarded([&] {
        if (columns == 0 || columns > 64 || cap == 0 || upload_us == 0 || compute_us == 0 || !sizes_out || !groups_out || !idle_us_out)
            throw ProveFailure{EZK_ERR_INVALID_ARGUMENT, "bad self-test arguments"};
        std::vector<uint64_t> ends;   // end time of every launch, in order
        std::vector<uint32_t> sizes;
        uint32_t launched = 0;
        uint64_t idle = 0, gpu_free = 0;
        auto decide = [&](uint64_t now, uint32_t arrived) {
            uint32_t head = 0;
            while (head < ends.size() && ends[head] <= now) head++;
            const uint32_t pending = (uint32_t)ends.size() - head;
            const uint32_t avail = arrived - launched;
            if (!host_group_ready(avail, pending, pending ? sizes[head] : 0, cap, arrived == columns)) return false;
            const uint32_t k = std::min(avail, cap);
            const uint64_t ready_at = (uint64_t)(launched + k) * upload_us;  // its last column is on the device
            const uint64_t start = std::max({now, ready_at, gpu_free});
            if (launched > 0 && start > gpu_free) idle += start - gpu_free;
            gpu_free = start + (uint64_t)k * compute_us;
            ends.push_back(gpu_free), sizes.push_back(k);
            launched += k;
            return true;
        };
        for (uint32_t c = 0; c < columns; c++) decide((uint64_t)(c + 1) * upload_us, c + 1);
        uint64_t now = (uint64_t)columns * upload_us;
        for (int guard = 0; launched < columns; guard++) {
            if (guard > 1000) throw ProveFailure{EZK_ERR_INTERNAL, "the launch rule stalls"};
            if (decide(now, columns)) continue;
            uint64_t next = UINT64_MAX;  // nothing to launch yet: the next decision comes when a launch ends
            for (uint64_t e : ends)
                if (e > now) next = std::min(next, e);
            if (next == UINT64_MAX) throw ProveFailure{EZK_ERR_INTERNAL, "the launch rule waits for nothing"};
            now = next;
        }
        uint32_t total = 0;
        for (size_t i = 0; i < sizes.size(); i++) {
            if (sizes[i] == 0 || sizes[i] > cap) throw ProveFailure{EZK_ERR_INTERNAL, "launch size out of range"};
            sizes_out[i] = sizes[i], total += sizes[i];
        }
        if (total != columns) throw ProveFailure{EZK_ERR_INTERNAL, "columns lost or launched twice"};
        *groups_out = (uint32_t)sizes.size();
        *idle_us_out = idle;
    });
}
int ezk_selftest_host_field(const void* a, const void* b, size_t n, void* out4n) {
    return guarded([&] {
        if (!a || !b || !out4n) throw ProveFailure{EZK_ERR_INVALID_ARGUMENT, "null argument"};
        const uint8_t *pa = (const uint8_t*)a, *pb = (const uint8_t*)b;
        uint8_t* po = (uint8_t*)out4n;
        for (size_t i = 0; i < n; i++) {
            const Fp x = fp_load(pa + 16 * i), y = fp_load(pb + 16 * i);
            if (x.v >= Fp::modulus() || y.v >= Fp::modulus()) throw ProveFailure{EZK_ERR_INVALID_ARGUMENT, "non-canonical element"};
            Fp r[4];
            host_field_products(x, y, r[0], r[1], r[2], r[3]);
            for (int k = 0; k < 4; k++) fp_store(po + 16 * (4 * i + k), r[k]);
        }
    });
}
// Host-only model of the sharded Merkle commitment: leaf digests spread over G ranks by row ownership, routed with the
// all-to-all rule of dist/shard_layout.h, subtrees + host-side top, then every node looked up through shard_node_home
// and every batch-proof node list checked against the unsplit tree.
int ezk_selftest_shard_layout(uint32_t world, uint32_t log_leaves, uint64_t seed) {
    return guarded([&] {
        const uint32_t G = world, glog = G == 8 ? 3 : G == 4 ? 2 : G == 2 ? 1 : 0;
        if ((1u << glog) != G || log_leaves < 2 * glog + 1 || log_leaves > 20) throw ProveFailure{EZK_ERR_INVALID_ARGUMENT, "bad shape"};
        const uint64_t leaves = 1ull << log_leaves, ll = leaves >> glog, chunk = ll >> glog;
        SplitMix64 rng(seed);
        std::vector<Hash32> leaf(leaves);
        for (auto& h : leaf)
            for (int i = 0; i < 4; i++) {
                const uint64_t v = rng.next();
                memcpy(h.data() + 8 * i, &v, 8);
            }
        // the unsplit tree
        std::vector<Hash32> full(2 * leaves);
        for (uint64_t i = 0; i < leaves; i++) full[leaves + i] = leaf[i];
        for (uint64_t k = leaves - 1; k >= 1; k--) full[k] = merge_digests(full[2 * k], full[2 * k + 1]);
        // packed digests of every rank, the all-to-all, the unpacking
        std::vector<std::vector<Hash32>> packed(G, std::vector<Hash32>(ll)), sub(G, std::vector<Hash32>(2 * ll));
        for (uint32_t r = 0; r < G; r++)
            for (uint64_t t = 0; t < ll; t++) packed[r][t] = leaf[(uint64_t)G * t + r];
        for (uint32_t src = 0; src < G; src++)
            for (uint64_t t = 0; t < ll; t++) {
                const uint32_t dst = shard_leaf_destination(leaves, glog, t);
                const uint64_t local = shard_local_leaf(G, src, t % chunk);
                if (dst >= G || local >= ll) throw ProveFailure{EZK_ERR_INTERNAL, "routing out of range"};
                sub[dst][ll + local] = packed[src][t];
            }
        for (uint32_t q = 0; q < G; q++) {
            for (uint64_t u = 0; u < ll; u++)
                if (sub[q][ll + u] != leaf[(uint64_t)q * ll + u]) throw ProveFailure{EZK_ERR_INTERNAL, "a leaf digest reached the wrong subtree slot"};
            for (uint64_t k = ll - 1; k >= 1; k--) sub[q][k] = merge_digests(sub[q][2 * k], sub[q][2 * k + 1]);
        }
        std::vector<Hash32> top(2 * G);
        for (uint32_t q = 0; q < G; q++) top[G + q] = sub[q][1];
        for (uint32_t k = G - 1; k >= 1; k--) top[k] = merge_digests(top[2 * k], top[2 * k + 1]);
        if (top[1] != full[1]) throw ProveFailure{EZK_ERR_INTERNAL, "root of the split tree differs"};
        auto fetch = [&](uint64_t k) {
            const NodeHome h = shard_node_home(G, glog, k);
            return h.owner < 0 ? top[h.index] : sub[h.owner][h.index];
        };
        for (uint64_t k = 1; k < 2 * leaves; k++)
            if (fetch(k) != full[k]) throw ProveFailure{EZK_ERR_INTERNAL, "node lookup through shard_node_home differs"};
        // authentication paths of random query sets
        for (int round = 0; round < 8; round++) {
            std::vector<uint64_t> pos;
            for (int q = 0; q < 32; q++) pos.push_back(rng.next() & (leaves - 1));
            std::sort(pos.begin(), pos.end());
            pos.erase(std::unique(pos.begin(), pos.end()), pos.end());
            for (auto& list : batch_proof_node_indices(leaves, pos))
                for (uint64_t k : list)
                    if (fetch(k) != full[k]) throw ProveFailure{EZK_ERR_INTERNAL, "authentication path node differs"};
        }
    });
}

void ezk_default_options(ezk_options* out) {
    if (!out) return;
    out->num_queries = 32, out->blowup_factor = 8, out->grinding_factor = 0, out->field_extension = 1;
    out->fri_folding_factor = 8, out->fri_remainder_max_degree = 127;
}

void ezk_get_wire_compat(ezk_wire_compat* out) {
    if (!out) return;
    const WireCompat& c = wire_compat();
    out->ood_interleaved = c.ood_interleaved, out->remainder_low_to_high = c.remainder_low_to_high;
    out->trace_info_aux_rands_byte = c.trace_info_aux_rands_byte, out->reserved = 0, out->first_nonce = c.first_nonce;
}
void ezk_set_wire_compat(const ezk_wire_compat* in) {
    WireCompat c;
    if (in) {
        c.ood_interleaved = in->ood_interleaved != 0, c.remainder_low_to_high = in->remainder_low_to_high != 0;
        c.trace_info_aux_rands_byte = in->trace_info_aux_rands_byte != 0, c.first_nonce = in->first_nonce;
    }
    wire_compat() = c;
}

int ezk_prover_create(int device, ezk_prover** out) {
    return guarded([&] {
        if (!out) throw ProveFailure{EZK_ERR_INVALID_ARGUMENT, "out pointer is null"};
        auto p = std::make_unique<ezk_prover>();
        p->impl = std::make_unique<GpuProver>(device);
        *out = p.release();
    });
}
void ezk_prover_destroy(ezk_prover* p) { delete p; }

int ezk_prover_prove(ezk_prover* p, const ezk_trace* trace, const ezk_public_inputs* pub, const ezk_options* opt,
                     uint8_t** proof, size_t* proof_len) {
    return guarded([&] {
        if (!p || !trace || !proof || !proof_len || !trace->columns) throw ProveFailure{EZK_ERR_INVALID_ARGUMENT, "null argument"};
        if (trace->width != EZK_TRACE_WIDTH) throw ProveFailure{EZK_ERR_INVALID_ARGUMENT, "trace width must be 28"};
        for (uint32_t c = 0; c < trace->width; c++)
            if (!trace->columns[c]) throw ProveFailure{EZK_ERR_INVALID_ARGUMENT, "null trace column"};
        std::lock_guard<std::mutex> lock(p->mu);
        auto bytes = p->impl->prove(trace->columns, nullptr, trace->length, to_public(pub), to_options(opt));
        export_proof(bytes, proof, proof_len);
    });
}

int ezk_prover_prove_ops(ezk_prover* p, const ezk_trace* trace, const ezk_op_list* ops, const ezk_public_inputs* pub,
                         const ezk_options* opt, uint8_t** proof, size_t* proof_len) {
    return guarded([&] {
        if (!p || !trace || !ops || !proof || !proof_len || !trace->columns) throw ProveFailure{EZK_ERR_INVALID_ARGUMENT, "null argument"};
        if (trace->width != EZK_TRACE_WIDTH) throw ProveFailure{EZK_ERR_INVALID_ARGUMENT, "trace width must be 28"};
        for (uint32_t c = 0; c < trace->width; c++)
            if (!trace->columns[c] && !is_op_column(c)) throw ProveFailure{EZK_ERR_INVALID_ARGUMENT, "null trace column"};
        OpList ol;
        ol.codes = ops->codes, ol.count = ops->count, ol.last_row = reinterpret_cast<const uint8_t*>(ops->last_row);
        std::lock_guard<std::mutex> lock(p->mu);
        auto bytes = p->impl->prove(trace->columns, nullptr, trace->length, to_public(pub), to_options(opt), &ol);
        export_proof(bytes, proof, proof_len);
    });
}

int ezk_prover_prove_device(ezk_prover* p, const void* d_trace, uint64_t length, const ezk_public_inputs* pub,
                            const ezk_options* opt, uint8_t** proof, size_t* proof_len) {
    return guarded([&] {
        if (!p || !d_trace || !proof || !proof_len) throw ProveFailure{EZK_ERR_INVALID_ARGUMENT, "null argument"};
        std::lock_guard<std::mutex> lock(p->mu);
        auto bytes = p->impl->prove(nullptr, static_cast<const uint4*>(d_trace), length, to_public(pub), to_options(opt));
        export_proof(bytes, proof, proof_len);
    });
}

int ezk_prover_verify(ezk_prover* p, const uint8_t* proof, size_t proof_len, const ezk_public_inputs* pub,
                      uint32_t min_conjectured_security) {
    return guarded([&] {
        if (!p || !proof || !pub) throw ProveFailure{EZK_ERR_INVALID_ARGUMENT, "null argument"};
        std::lock_guard<std::mutex> lock(p->mu);
        p->impl->verify(proof, proof_len, to_public(pub), min_conjectured_security);
    });
}

int ezk_prove(const ezk_trace* trace, const ezk_public_inputs* pub, const ezk_options* opt, uint8_t** proof,
              size_t* proof_len) {
    {
        std::lock_guard<std::mutex> lock(g_default_mu);
        if (!g_default) {
            int rc = ezk_prover_create(0, &g_default);
            if (rc != EZK_OK) return rc;
        }
    }
    return ezk_prover_prove(g_default, trace, pub, opt, proof, proof_len);
}

int ezk_prover_stage_times(const ezk_prover* p, float* ms_out) {
    return guarded([&] {
        if (!p || !ms_out) throw ProveFailure{EZK_ERR_INVALID_ARGUMENT, "null argument"};
        memcpy(ms_out, p->impl->stage_ms(), EZK_STAGE_COUNT * sizeof(float));
    });
}

int ezk_comm_unique_id(uint8_t out[128]) {
    return guarded([&] {
        if (!out) throw ProveFailure{EZK_ERR_INVALID_ARGUMENT, "null argument"};
        Comm::unique_id(out);
    });
}
int ezk_prover_join(ezk_prover* p, int rank, int world, const uint8_t unique_id[128]) {
    return guarded([&] {
        if (!p || !unique_id) throw ProveFailure{EZK_ERR_INVALID_ARGUMENT, "null argument"};
        std::lock_guard<std::mutex> lock(p->mu);
        p->impl->join(rank, world, unique_id);
    });
}

int ezk_local_group_create(int world, ezk_group** out) {
    return guarded([&] {
        if (!out) throw ProveFailure{EZK_ERR_INVALID_ARGUMENT, "null argument"};
        auto g = std::make_unique<ezk_group>();
        g->impl = std::make_unique<LocalGroup>(world);
        *out = g.release();
    });
}
void ezk_local_group_destroy(ezk_group* g) { delete g; }
int ezk_prover_join_local(ezk_prover* p, ezk_group* g, int rank) {
    return guarded([&] {
        if (!p) throw ProveFailure{EZK_ERR_INVALID_ARGUMENT, "null argument"};
        std::lock_guard<std::mutex> lock(p->mu);
        p->impl->join_local(rank, g ? g->impl.get() : nullptr);
    });
}
int ezk_prover_timer_start(ezk_prover* p) {
    return guarded([&] {
        if (!p) throw ProveFailure{EZK_ERR_INVALID_ARGUMENT, "null argument"};
        p->impl->timer_start();
    });
}
int ezk_prover_timer_stop(ezk_prover* p, float* ms_out) {
    return guarded([&] {
        if (!p || !ms_out) throw ProveFailure{EZK_ERR_INVALID_ARGUMENT, "null argument"};
        *ms_out = p->impl->timer_stop();
    });
}
void ezk_profile_enable(int on) { profile_enable(on != 0); }
void ezk_profile_reset(void) { profile_reset(); }
int ezk_profile_kernel_count(void) { return K_COUNT; }
const char* ezk_profile_kernel_name(int kernel) { return kernel_name(kernel); }
void ezk_profile_read(int kernel, uint64_t* launches, double* ms, uint64_t* algo_bytes) {
    profile_read(kernel, launches, ms, algo_bytes);
}

int ezk_prover_artifact(ezk_prover* p, int which, void* dst, size_t cap, size_t* size_out) {
    return guarded([&] {
        if (!p) throw ProveFailure{EZK_ERR_INVALID_ARGUMENT, "null argument"};
        std::lock_guard<std::mutex> lock(p->mu);
        auto bytes = p->impl->artifact(which);
        if (size_out) *size_out = bytes.size();
        if (dst) memcpy(dst, bytes.data(), std::min(cap, bytes.size()));
    });
}

int ezk_stage_lde(ezk_prover* p, const void* columns, uint32_t width, uint64_t n, void* lde_out) {
    return guarded([&] {
        if (!p || !columns || !lde_out) throw ProveFailure{EZK_ERR_INVALID_ARGUMENT, "null argument"};
        std::lock_guard<std::mutex> lock(p->mu);
        p->impl->stage_lde(columns, width, n, lde_out);
    });
}
int ezk_stage_ntt(ezk_prover* p, const void* columns, uint32_t width, uint64_t n, int inverse, void* out) {
    return guarded([&] {
        if (!p || !columns || !out) throw ProveFailure{EZK_ERR_INVALID_ARGUMENT, "null argument"};
        std::lock_guard<std::mutex> lock(p->mu);
        p->impl->stage_ntt(columns, width, n, inverse != 0, out);
    });
}
int ezk_stage_merkle(ezk_prover* p, const void* table, uint32_t width, uint64_t rows, void* nodes_out) {
    return guarded([&] {
        if (!p || !table || !nodes_out) throw ProveFailure{EZK_ERR_INVALID_ARGUMENT, "null argument"};
        std::lock_guard<std::mutex> lock(p->mu);
        p->impl->stage_merkle(table, width, rows, nodes_out);
    });
}
int ezk_stage_fri_fold(ezk_prover* p, const void* evals, uint64_t s, const void* alpha16, void* next_out) {
    return guarded([&] {
        if (!p || !evals || !alpha16 || !next_out) throw ProveFailure{EZK_ERR_INVALID_ARGUMENT, "null argument"};
        std::lock_guard<std::mutex> lock(p->mu);
        p->impl->stage_fri_fold(evals, s, fp_load(static_cast<const uint8_t*>(alpha16)), next_out);
    });
}
int ezk_stage_eval_frames(ezk_prover* p, const void* cur, const void* next, const void* periodic, uint32_t nframes,
                          uint32_t lwe_delta, void* out20) {
    return guarded([&] {
        if (!p || !cur || !next || !periodic || !out20 || !nframes) throw ProveFailure{EZK_ERR_INVALID_ARGUMENT, "null argument"};
        std::lock_guard<std::mutex> lock(p->mu);
        p->impl->stage_eval_frames(cur, next, periodic, nframes, lwe_delta, out20);
    });
}
int ezk_stage_eval_frames_sum(ezk_prover* p, const void* cur, const void* next, const void* periodic, uint32_t nframes,
                              uint32_t lwe_delta, const void* tcoef20, void* out1) {
    return guarded([&] {
        if (!p || !cur || !next || !periodic || !tcoef20 || !out1) throw ProveFailure{EZK_ERR_INVALID_ARGUMENT, "null argument"};
        Fp tc[20];
        for (int j = 0; j < 20; j++) tc[j] = fp_load(static_cast<const uint8_t*>(tcoef20) + 16 * j);
        std::lock_guard<std::mutex> lock(p->mu);
        p->impl->stage_eval_frames_sum(cur, next, periodic, nframes, lwe_delta, tc, out1);
    });
}
int ezk_bench_lde_merkle(ezk_prover* p, uint32_t width, uint64_t n, int iters, float* lde_ms, float* merkle_ms) {
    return guarded([&] {
        if (!p || !lde_ms || !merkle_ms) throw ProveFailure{EZK_ERR_INVALID_ARGUMENT, "null argument"};
        std::lock_guard<std::mutex> lock(p->mu);
        p->impl->bench_lde_merkle(width, n, iters, lde_ms, merkle_ms);
    });
}
int ezk_bench_fri(ezk_prover* p, uint64_t n, int iters, float* fri_ms) {
    return guarded([&] {
        if (!p || !fri_ms) throw ProveFailure{EZK_ERR_INVALID_ARGUMENT, "null argument"};
        std::lock_guard<std::mutex> lock(p->mu);
        p->impl->bench_fri(n, iters, fri_ms);
    });
}

// ---- host VM ----
int ezk_program_compile(const char* source, ezk_program** out) {
    return guarded([&] {
        if (!source || !out) throw ProveFailure{EZK_ERR_INVALID_ARGUMENT, "null argument"};
        auto p = std::make_unique<ezk_program>();
        p->prog = Program::compile(source);
        *out = p.release();
    });
}
void ezk_program_free(ezk_program* p) { delete p; }
size_t ezk_program_len(const ezk_program* p) { return p ? p->prog.code.size() : 0; }
void ezk_program_ops(const ezk_program* p, uint8_t* codes, uint8_t* values) {
    if (!p) return;
    for (size_t i = 0; i < p->prog.code.size(); i++) {
        if (codes) codes[i] = p->prog.code[i].code;
        if (values) values[i] = p->prog.code[i].value;
    }
}
void ezk_program_hash(const ezk_program* p, uint8_t out[2][16]) {
    if (!p) return;
    fp_store(out[0], p->prog.hash[0]);
    fp_store(out[1], p->prog.hash[1]);
}
size_t ezk_program_display(const ezk_program* p, char* dst, size_t cap) {
    if (!p) return 0;
    std::string s = p->prog.to_string();
    if (dst && cap) {
        size_t k = std::min(cap - 1, s.size());
        memcpy(dst, s.data(), k);
        dst[k] = 0;
    }
    return s.size() + 1;
}

int ezk_vm_execute(const ezk_program* prog, const uint8_t* public_tape, size_t public_len, const void* secret_elems,
                   size_t num_ciphertexts, uint32_t lwe_k, uint32_t lwe_delta, uint64_t last_row_seed,
                   ezk_execution** out) {
    return guarded([&] {
        if (!prog || !out || (public_len && !public_tape) || (num_ciphertexts && !secret_elems))
            throw ProveFailure{EZK_ERR_INVALID_ARGUMENT, "null argument"};
        LweParams lwe;
        lwe.k = lwe_k, lwe.delta = lwe_delta;
        std::vector<uint8_t> pub(public_tape, public_tape + public_len);
        std::vector<Fp> secret(num_ciphertexts * lwe.lwe_size());
        for (size_t i = 0; i < secret.size(); i++) secret[i] = fp_load(static_cast<const uint8_t*>(secret_elems) + 16 * i);
        auto e = std::make_unique<ezk_execution>();
        e->trace = execute(prog->prog, pub, secret, lwe, last_row_seed);
        *out = e.release();
    });
}
void ezk_execution_free(ezk_execution* e) { delete e; }
uint64_t ezk_execution_length(const ezk_execution* e) { return e ? e->trace.n : 0; }
const uint8_t* ezk_execution_column(const ezk_execution* e, uint32_t c) {
    if (!e || c >= e->trace.columns.size()) return nullptr;
    return reinterpret_cast<const uint8_t*>(e->trace.columns[c].data());
}
void ezk_execution_outputs(const ezk_execution* e, uint8_t out[16][16]) {
    if (!e) return;
    for (int i = 0; i < 16; i++) fp_store(out[i], e->trace.outputs[i]);
}

int ezk_synthetic_case(int kind, uint32_t log_n, uint32_t lwe_k, uint32_t lwe_delta, uint64_t seed, ezk_program** prog,
                       ezk_execution** exec) {
    return guarded([&] {
        if (!prog || !exec || kind < 1 || kind > 3 || log_n < 7 || log_n > 26)
            throw ProveFailure{EZK_ERR_INVALID_ARGUMENT, "bad synthetic case parameters"};
        LweParams lwe;
        lwe.k = lwe_k, lwe.delta = lwe_delta;
        SyntheticCase c = make_synthetic(kind, log_n, lwe, seed);
        auto e = std::make_unique<ezk_execution>();
        e->trace = execute(c.program, c.pub, c.secret, lwe, seed ^ 0x5EEDULL);
        auto p = std::make_unique<ezk_program>();
        p->prog = std::move(c.program);
        *prog = p.release();
        *exec = e.release();
    });
}

void ezk_lwe_keygen(uint32_t k, uint64_t seed, void* key_out) {
    LweParams p;
    p.k = k;
    LweKey key = lwe_keygen(p, seed);
    for (uint32_t i = 0; i < k; i++) fp_store(static_cast<uint8_t*>(key_out) + 16 * i, key.key[i]);
}
void ezk_lwe_encrypt(const void* key, uint32_t k, uint32_t delta, double std_dev, uint8_t value, uint64_t seed,
                     void* ct_out) {
    LweKey lk;
    lk.params.k = k, lk.params.delta = delta, lk.params.std_dev = std_dev;
    for (uint32_t i = 0; i < k; i++) lk.key.push_back(fp_load(static_cast<const uint8_t*>(key) + 16 * i));
    auto ct = lwe_encrypt(lk, value, seed);
    for (uint32_t i = 0; i <= k; i++) fp_store(static_cast<uint8_t*>(ct_out) + 16 * i, ct[i]);
}
uint8_t ezk_lwe_decrypt(const void* key, uint32_t k, uint32_t delta, const void* ct) {
    LweKey lk;
    lk.params.k = k, lk.params.delta = delta;
    for (uint32_t i = 0; i < k; i++) lk.key.push_back(fp_load(static_cast<const uint8_t*>(key) + 16 * i));
    std::vector<Fp> c(k + 1);
    for (uint32_t i = 0; i <= k; i++) c[i] = fp_load(static_cast<const uint8_t*>(ct) + 16 * i);
    return lwe_decrypt(lk, c.data());
}

}  // extern "C"
