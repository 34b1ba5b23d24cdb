/* C ABI of the B200 proving backend for Encrypt-zkVM.
 *
 * Drop-in boundary: `prover.prove(trace)` at vm/src/lib.rs:26, i.e. the `winterfell::Prover` implementation
 * `ExecutionProver` of prover/src/lib.rs:17-77 (BaseField f128, ProcessorAir, Blake3_256, DefaultRandomCoin,
 * DefaultTraceLde, DefaultConstraintEvaluator).  The reference has no FFI of its own (pure Rust, no unsafe);
 * these are the entry points a Rust shim would bind (see INTEGRATION.md for the `extern "C"` block).
 *
 * All field elements cross the boundary as 16 little-endian bytes of the canonical u128 value - exactly the
 * memory of `winterfell::math::fields::f128::BaseElement`, so a `&[BaseElement]` can be passed as a pointer.
 * Every function returns EZK_OK (0) or a negative error code; ezk_last_error() gives the message (thread-local).
 * The library never falls back to the CPU: without a CUDA device every compute entry point fails with
 * EZK_ERR_NO_DEVICE.
 */
#ifndef EZKVM_PROVER_H
#define EZKVM_PROVER_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EZK_TRACE_WIDTH 28 /* vm/src/processor/mod.rs:76-84 */

enum ezk_status {
    EZK_OK = 0,
    EZK_ERR_INVALID_ARGUMENT = -1,
    EZK_ERR_UNSUPPORTED_FIELD_EXTENSION = -2, /* ProverError::UnsupportedFieldExtension */
    EZK_ERR_CONSTRAINT_DEGREE = -3,           /* ProverError::MismatchedConstraintPolynomialDegree: trace violates the AIR */
    EZK_ERR_DEEP_DEGREE = -4,                 /* winterfell's assert on the DEEP polynomial degree / FRI remainder */
    EZK_ERR_NO_DEVICE = -5,
    EZK_ERR_CUDA = -6,
    EZK_ERR_VM = -7,                          /* ProgramError / ProcessorError (message mirrors the reference's Display) */
    EZK_ERR_INTERNAL = -8,
    EZK_ERR_VERIFICATION = -9                 /* VerifierError: the proof was rejected (ezk_last_error() says why) */
};

/* winterfell::ProofOptions::new(32, 8, 0, FieldExtension::None, 8, 127) - vm/src/lib.rs:20 */
typedef struct ezk_options {
    uint32_t num_queries;              /* 32 */
    uint32_t blowup_factor;            /* 8 (the only supported value) */
    uint32_t grinding_factor;          /* 0 */
    uint32_t field_extension;          /* 1 = FieldExtension::None (the only supported value) */
    uint32_t fri_folding_factor;       /* 8 (the only supported value) */
    uint32_t fri_remainder_max_degree; /* 127 */
} ezk_options;

/* air::PublicInputs (air/src/lib.rs:18-47) + the two ServerKey parameters the AIR reads
 * (fhe/src/parameters.rs:4-21, used at air/src/constrains.rs:113,129-130,151). */
typedef struct ezk_public_inputs {
    uint8_t program_hash[2][16];
    uint8_t stack_outputs[16][16];
    uint32_t lwe_k;     /* must be 4: the AIR hard-codes lwe_size 5 (constrains.rs:104-105,175) */
    uint32_t lwe_delta; /* ciphertext_modulus / plaintext_modulus */
} ezk_public_inputs;

/* TraceTable<BaseElement> (vm/src/lib.rs:18): `width` column pointers, each `length` elements of 16 bytes. */
typedef struct ezk_trace {
    const uint8_t* const* columns;
    uint32_t width;  /* 28 */
    uint64_t length; /* power of two, >= 64 */
} ezk_trace;

typedef struct ezk_prover ezk_prover; /* holds the CUDA stream, twiddle tables and a reusable device workspace */

/* ---- library ---- */
const char* ezk_last_error(void);
const char* ezk_version(void);
int ezk_device_count(void);           /* number of visible CUDA devices (0 without a GPU; never fails) */
uint64_t ezk_kernel_launch_count(void); /* kernels launched by this library since load */
void ezk_free(void* p);               /* frees buffers returned by this library */
void ezk_default_options(ezk_options* out);

/* Byte-level details of the winterfell 0.9.0 proof format that the reference tree cannot settle (the engine is an
 * un-vendored dependency).  The proof writer and the verifier both read them from one
 * process-wide copy of this struct; the defaults are the current reading of the 0.9.0 sources.  Set it before
 * proving / verifying (not while a proof is in flight).  Mirrors `Compat` of oracle/stark.hpp field by field. */
typedef struct ezk_wire_compat {
    uint32_t ood_interleaved;           /* 1: OOD trace states as [cur_0, next_0, cur_1, next_1, ...] (default) */
    uint32_t remainder_low_to_high;     /* 1: FRI remainder coefficients, constant term first (default) */
    uint32_t trace_info_aux_rands_byte; /* 1: TraceInfo = u8 width, u8 aux width, u8 aux rands, u8 log2 n, u16 meta (default) */
    uint32_t reserved;
    uint64_t first_nonce;               /* where the grinding search starts (default 1) */
} ezk_wire_compat;
void ezk_get_wire_compat(ezk_wire_compat* out);
void ezk_set_wire_compat(const ezk_wire_compat* in); /* NULL restores the defaults */

/* Host-side self-test of the threaded copy used by the staged trace upload (EZK_STAGED_UPLOAD=1):
 * copies `bytes` pseudo-random bytes with `threads` threads in chunks of every alignment class and compares.
 * Needs no GPU.  Returns EZK_OK or EZK_ERR_INTERNAL. */
int ezk_selftest_copy_pool(uint32_t threads, size_t bytes);

/* Host-side self-test of the index arithmetic of the sharded proof (csrc/dist/shard_layout.h): models the leaf-digest
 * all-to-all, the per-rank subtrees and the host-side top levels of a split Merkle commitment with `world` ranks and
 * 2^log_leaves leaves, and checks root, every node lookup and random authentication paths against the unsplit tree.
 * Needs no GPU.  Returns EZK_OK or EZK_ERR_INTERNAL. */
int ezk_selftest_shard_layout(uint32_t world, uint32_t log_leaves, uint64_t seed);

/* Host-side replay of the rule that sizes the interpolation + LDE launches of a host trace (csrc/host/launch_groups.h):
 * `columns` columns arriving every upload_us, a launch of k columns taking k * compute_us, at most `cap` columns per
 * launch.  Writes the launch sizes (up to `columns` entries), their number and the time the compute stream sat idle
 * after its first launch.  Needs no GPU.  Returns EZK_OK, EZK_ERR_INVALID_ARGUMENT or EZK_ERR_INTERNAL (a column lost). */
int ezk_selftest_launch_groups(uint32_t columns, uint32_t cap, uint32_t upload_us, uint32_t compute_us, uint32_t* sizes_out,
                               uint32_t* groups_out, uint64_t* idle_us_out);

/* Host-side self-test of the f128 arithmetic behind the transcript and the VM: for n pairs (a_i, b_i) of canonical
 * elements writes a_i * b_i (portable product), a_i * b_i (the product the Rescue sponge uses), a_i^2 and
 * a_i^INV_ALPHA (the sponge's addition chain) - 4 * n elements - for the caller to compare with big integers.
 * Needs no GPU. */
int ezk_selftest_host_field(const void* a, const void* b, size_t n, void* out4n);

/* ---- prover object: ExecutionProver::new (prover/src/lib.rs:25-37) ---- */
int ezk_prover_create(int device, ezk_prover** out);
void ezk_prover_destroy(ezk_prover* p);

/* ExecutionProver::prove (prover/src/lib.rs:40-77 via winterfell::Prover::prove): host trace in, serialized
 * `winterfell::Proof` bytes out (`Proof::to_bytes` layout).  *proof is malloc'ed; release with ezk_free. */
int ezk_prover_prove(ezk_prover* p, const ezk_trace* trace, const ezk_public_inputs* pub, const ezk_options* opt,
                     uint8_t** proof, size_t* proof_len);
/* Same, with the trace already resident in device memory as 28 contiguous columns (column c at
 * d_trace + c * length * 16).  The buffer is not modified. */
int ezk_prover_prove_device(ezk_prover* p, const void* d_trace, uint64_t length, const ezk_public_inputs* pub,
                            const ezk_options* opt, uint8_t** proof, size_t* proof_len);
/* One-shot convenience with a lazily created default prover on device 0 (what a Rust shim calls). */
int ezk_prove(const ezk_trace* trace, const ezk_public_inputs* pub, const ezk_options* opt, uint8_t** proof,
              size_t* proof_len);

/* The same proof from a trace whose "bookkeeping" columns are built on the device (scoped to what is
 * parallel): clk (column 0), the five op bits (1..5), the chiplet flag (6) and the stack depth (11) are pure functions
 * of the executed operation list - one code byte per operation (vm/src/processor/opcodes.rs:30-43, in execution order,
 * the compiler's padding NOOPs included: `Program::get_code()`) - so a caller may pass NULL for those eight column
 * pointers and the list instead: 8 x 16 bytes per row less to upload.  The Rescue sponge columns (7..10) and the
 * stack registers (12..27) still come from the host VM (vm/src/processor/chiplets.rs:92-112, stack.rs:48-70).
 * last_row: the 28 values of row length-1 (the reference's thread_rng row, vm/src/processor/mod.rs:86-92).
 * Returns the bytes ezk_prover_prove returns for the full trace. */
typedef struct ezk_op_list {
    const uint8_t* codes;          /* count operation codes */
    uint64_t count;                /* executed operations, < trace length */
    const uint8_t (*last_row)[16]; /* 28 elements */
} ezk_op_list;
int ezk_prover_prove_ops(ezk_prover* p, const ezk_trace* trace, const ezk_op_list* ops, const ezk_public_inputs* public_inputs,
                         const ezk_options* options, uint8_t** proof, size_t* proof_len);

/* winterfell::verify::<ProcessorAir, Blake3_256, DefaultRandomCoin<Blake3_256>>(proof, pub_inputs,
 * &AcceptableOptions::MinConjecturedSecurity(min_conjectured_security)) as called at vm/src/lib.rs:91-98 and
 * examples/linear_regression/src/main.rs:81-85.  Returns EZK_OK when the proof is accepted and
 * EZK_ERR_VERIFICATION when it is rejected.  Transcript / Merkle / FRI checks run on the host; the AIR's
 * transition constraints at the out-of-domain point run on the GPU (same code as the prover). */
int ezk_prover_verify(ezk_prover* p, const uint8_t* proof, size_t proof_len, const ezk_public_inputs* pub,
                      uint32_t min_conjectured_security);

/* ---- one proof sharded over the GPUs of a box (no reference counterpart: the reference is one thread) ----
 * One process per GPU.  Rank 0 calls ezk_comm_unique_id and hands the 128 bytes to the other ranks (any
 * transport; the Python binding uses torch.distributed); every rank then calls ezk_prover_join on its own
 * prover (a collective: it builds the NCCL communicator).  From then on ezk_prover_prove / _prove_device must be
 * called by ALL ranks with the same trace length / public inputs / options.  Rank r of `world`
 *   - reads only the trace columns c with c mod world = r (the other column pointers / device columns are never
 *     touched: a caller may hand every rank only its own columns), interpolates them and all-gathers the coefficients;
 *   - extends, hashes and evaluates the LDE rows i with i mod world = r (whole cosets);
 *   - holds the Merkle subtree over the leaves [r L / world, (r + 1) L / world) of every commitment: leaf digests move
 *     with one all-to-all, only the `world` subtree roots are gathered;
 *   - interpolates the constraint evaluations on its cosets; the composition columns come from an all-gather of those
 *     coefficient arrays and an 8-point inverse DFT across the cosets;
 *   - evaluates the out-of-domain frame for its columns, the DEEP composition and the large FRI layers for its rows.
 * Every rank returns the same proof bytes, identical to the single-GPU proof.  world in {1, 2, 4, 8}; world = 1
 * leaves the group.  A proof that fails input validation fails on every rank; any other failure on one rank of an
 * NCCL group leaves the others waiting in a collective (abort the processes). */
int ezk_comm_unique_id(uint8_t out[128]);
int ezk_prover_join(ezk_prover* p, int rank, int world, const uint8_t unique_id[128]);

/* The same sharded proof between several provers of ONE process (one host thread per prover, any mix of devices,
 * also all on one GPU): the exchanges are host-synchronised device copies instead of NCCL.  Not a performance path:
 * it runs the whole sharded pipeline on a one-GPU box (tests).  group = NULL leaves the group. */
typedef struct ezk_group ezk_group;
int ezk_local_group_create(int world, ezk_group** out);
void ezk_local_group_destroy(ezk_group* g);
int ezk_prover_join_local(ezk_prover* p, ezk_group* g, int rank);

/* Device-time of the stages of the last proof, milliseconds (CUDA events on the prover's stream). */
enum ezk_stage {
    EZK_STAGE_UPLOAD = 0,     /* host -> device trace copy */
    EZK_STAGE_TRACE_LDE,      /* 28 x iNTT + coset LDE          (DefaultTraceLde::new) */
    EZK_STAGE_TRACE_COMMIT,   /* BLAKE3 rows + Merkle tree */
    EZK_STAGE_CONSTRAINTS,    /* divisor inverses + fused AIR evaluation (DefaultConstraintEvaluator) */
    EZK_STAGE_COMPOSITION,    /* iNTT(8n) + 7 LDEs + commitment */
    EZK_STAGE_DEEP,           /* OOD evaluation + DEEP composition + LDE */
    EZK_STAGE_FRI,            /* layers + remainder */
    EZK_STAGE_QUERIES,        /* openings + proof assembly */
    EZK_STAGE_COUNT
};
int ezk_prover_stage_times(const ezk_prover* p, float* ms_out /* EZK_STAGE_COUNT */);

/* Brackets any sequence of calls with CUDA events on the prover's stream (device-side wall time, includes the
 * host transcript gaps between kernels): start, run, stop -> milliseconds. */
int ezk_prover_timer_start(ezk_prover* p);
int ezk_prover_timer_stop(ezk_prover* p, float* ms_out);

/* Per-kernel profile (CUDA events around every launch; off by default). */
void ezk_profile_enable(int on);
void ezk_profile_reset(void);
int ezk_profile_kernel_count(void);
const char* ezk_profile_kernel_name(int kernel);
/* totals since the last reset: launches, summed device milliseconds, summed algorithmic bytes */
void ezk_profile_read(int kernel, uint64_t* launches, double* ms, uint64_t* algo_bytes);

/* Intermediate values of the last proof, for stage-by-stage parity tests. Copies up to `cap` bytes into dst and
 * returns the full size in *size_out (call with dst = NULL to query). Elements are 16 LE bytes, digests 32. */
enum ezk_artifact {
    EZK_ART_TRACE_ROOT = 0,
    EZK_ART_CONSTRAINT_ROOT,
    EZK_ART_COMBINED,         /* L constraint evaluations */
    EZK_ART_OOD_TRACE,        /* 56 elements, interleaved cur/next */
    EZK_ART_OOD_CONSTRAINTS,  /* 7 elements */
    EZK_ART_DEEP_EVALS,       /* L elements */
    EZK_ART_FRI_ROOTS,        /* (layers + 1) digests, last = remainder commitment */
    EZK_ART_REMAINDER,        /* remainder coefficients */
    EZK_ART_POSITIONS,        /* u64 query positions (sorted, unique) */
    EZK_ART_TRACE_LDE,        /* column-major 28 x L */
    EZK_ART_CONSTRAINT_LDE,   /* column-major 7 x L */
    EZK_ART_TRACE_POLYS,      /* column-major 28 x n coefficients, scaled by 3^m */
    EZK_ART_COUNT
};
int ezk_prover_artifact(ezk_prover* p, int which, void* dst, size_t cap, size_t* size_out);

/* ---- stage-level entry points (host buffers in/out; used by parity tests and the stage sweep) ---- */
/* Column-major `width` x n values -> column-major `width` x 8n LDE over the coset 3*<w_8n> (iNTT + coset NTT). */
int ezk_stage_lde(ezk_prover* p, const void* columns, uint32_t width, uint64_t n, void* lde_out);
/* Column-major `width` x rows table -> 2*rows digests (node k at 32*k; node 1 = root; leaves from `rows`). */
int ezk_stage_merkle(ezk_prover* p, const void* table, uint32_t width, uint64_t rows, void* nodes_out);
/* FRI over evaluations on 3*<w_s>: returns (layers+1) roots, the remainder coefficients and the folded layers.
 * alphas come from the caller (one per layer) so the stage is testable without a transcript. */
int ezk_stage_fri_fold(ezk_prover* p, const void* evals, uint64_t s, const void* alpha16, void* next_out);
/* 20 transition-constraint values for explicit frames (reference unit tests: air/src/tests/mod.rs). */
int ezk_stage_eval_frames(ezk_prover* p, const void* cur, const void* next, const void* periodic, uint32_t nframes,
                          uint32_t lwe_delta, void* out20);
/* The same frames through the PRODUCTION path of the constraint kernel (selector-grouped accumulation, flagged
 * arithmetic with its exact redo): out1[f] = sum_j tcoef20[j] * r_j(frame f), one element per frame. */
int ezk_stage_eval_frames_sum(ezk_prover* p, const void* cur, const void* next, const void* periodic, uint32_t nframes,
                              uint32_t lwe_delta, const void* tcoef20, void* out1);
/* Plain transform of column-major `width` x n values; inverse != 0 -> interpolation (scaled by 1/n). */
int ezk_stage_ntt(ezk_prover* p, const void* columns, uint32_t width, uint64_t n, int inverse, void* out);
/* Device-resident stage benchmarks for the LDE/Merkle/FRI sweep: run `iters` times on synthetic device data,
 * return the average device time per iteration in milliseconds. */
int ezk_bench_lde_merkle(ezk_prover* p, uint32_t width, uint64_t n, int iters, float* lde_ms, float* merkle_ms);
int ezk_bench_fri(ezk_prover* p, uint64_t n, int iters, float* fri_ms);

/* ---- host VM (vm crate restated; stays on the host per the north star) ---- */
typedef struct ezk_program ezk_program;
typedef struct ezk_execution ezk_execution;

/* Program::compile (vm/src/program/mod.rs:37-96). On error returns EZK_ERR_VM and ezk_last_error() holds
 * "program error at {step}: {message}" exactly as the reference formats it. */
int ezk_program_compile(const char* source, ezk_program** out);
void ezk_program_free(ezk_program* p);
size_t ezk_program_len(const ezk_program* p);
void ezk_program_ops(const ezk_program* p, uint8_t* codes, uint8_t* values);
void ezk_program_hash(const ezk_program* p, uint8_t out[2][16]);
/* Display impl: "push(1) noop ..." ; returns bytes needed (incl. NUL), writes at most cap. */
size_t ezk_program_display(const ezk_program* p, char* dst, size_t cap);

/* Processor::run + Processor::trace (vm/src/processor/mod.rs:61-95). secret: num_ciphertexts * (lwe_k+1) elements.
 * The last trace row is filled from SplitMix64(last_row_seed) instead of thread_rng. */
int ezk_vm_execute(const ezk_program* prog, const uint8_t* public_tape, size_t public_len, const void* secret_elems,
                   size_t num_ciphertexts, uint32_t lwe_k, uint32_t lwe_delta, uint64_t last_row_seed,
                   ezk_execution** out);
void ezk_execution_free(ezk_execution* e);
uint64_t ezk_execution_length(const ezk_execution* e);
const uint8_t* ezk_execution_column(const ezk_execution* e, uint32_t c); /* length * 16 bytes */
void ezk_execution_outputs(const ezk_execution* e, uint8_t out[16][16]);

/* Synthetic benchmark programs: kind 1 scalar, 2 ciphertext, 3 mixed. Produces the
 * program, its tapes and the executed trace of length 2^log_n. */
int ezk_synthetic_case(int kind, uint32_t log_n, uint32_t lwe_k, uint32_t lwe_delta, uint64_t seed, ezk_program** prog,
                       ezk_execution** exec);

/* LWE client side with a seeded PRNG (fhe/src/server_key.rs:19-76): key = k elements, ciphertext = k+1. */
void ezk_lwe_keygen(uint32_t k, uint64_t seed, void* key_out);
void ezk_lwe_encrypt(const void* key, uint32_t k, uint32_t delta, double std_dev, uint8_t value, uint64_t seed,
                     void* ct_out);
uint8_t ezk_lwe_decrypt(const void* key, uint32_t k, uint32_t delta, const void* ct);

#ifdef __cplusplus
}
#endif
#endif /* EZKVM_PROVER_H */
