// Host-side VM of the reference (north star: "the vm crate still executes the program and builds
// the System/Stack/Decoder/Chiplets execution trace on the host").  The reference's host is Rust;
// this environment has no Rust toolchain, so the same behaviour is provided in C++ above the C ABI:
//   Program::compile           vm/src/program/mod.rs:37-96        (text -> padded op list + Rescue hash)
//   Rescue128 sponge           crypto/src/rescue.rs:16-118
//   Processor::run / trace     vm/src/processor/mod.rs:61-116      (28 columns, n = next_pow2(capacity+1))
//   ServerKey LWE ops          fhe/src/server_key.rs:19-124
// Error strings mirror vm/src/program/errors.rs and vm/src/processor/errors.rs.
#pragma once
#include "../field/f128_host.h"
#include <cstdint>
#include <string>
#include <vector>

namespace ezk {

enum OpCode : uint8_t {  // vm/src/processor/opcodes.rs:30-43
    OP_NOOP = 0b00000,
    OP_PUSH = 0b10000,
    OP_READ = 0b10001,
    OP_READ2 = 0b10010,
    OP_ADD = 0b01000,
    OP_MUL = 0b01001,
    OP_SADD = 0b01010,
    OP_SMUL = 0b01100,
    OP_ADD2 = 0b01011,
};

struct Operation {
    uint8_t code;
    uint8_t value;   // push immediate, 0 otherwise (opcodes.rs:92-99)
    std::string to_string() const;
};

struct VmError {
    std::string message;  // already formatted like the reference's Display impls
};

struct RescueSponge {  // crypto/src/rescue.rs:16-60
    Fp state[4];
    size_t step = 0;
    void update(uint8_t op_code, uint8_t op_value);
};

struct Program {
    std::vector<Operation> code;
    Fp hash[2];
    // Sponge state after each operation (4 elements per op), kept from the hashing pass of compile():
    // the chiplet columns of the trace are this same chain (chiplets.rs:92-112), and it is the one
    // inherently sequential, expensive part of the VM (one 128-bit exponentiation per lane per op), so
    // execute() copies it instead of running the chain a second time.  Empty => execute() recomputes.
    std::vector<Fp> sponge_states;
    // throws VmError("program error at {step}: {message}")
    static Program compile(const std::string& source);
    std::string to_string() const;  // "push(1) noop ..." (program/mod.rs:128-138)
};

struct LweParams {  // fhe/src/parameters.rs
    uint32_t plaintext_modulus = 8, ciphertext_modulus = 128, delta = 16;
    uint32_t k = 4;
    double std_dev = 2.412390240121573e-5;
    uint32_t lwe_size() const { return k + 1; }
};

struct ExecutionTrace {
    size_t n = 0;
    std::vector<std::vector<Fp>> columns;  // 28 x n  (vm/src/processor/mod.rs:76-84)
    Fp outputs[16];                        // Processor::output (mod.rs:97-101)
};

// Runs the program (Processor::run) and assembles the trace (Processor::trace). `secret` holds the
// ciphertext tape, lwe_size elements per ciphertext.  The reference overwrites the last row with
// thread_rng values in [1, 2^128) (mod.rs:86-92); here they come from SplitMix64(last_row_seed) so
// runs are reproducible.  Throws VmError with the reference's "stack error at ..." / "chiplets error at ..." text.
ExecutionTrace execute(const Program& program, const std::vector<uint8_t>& pub, const std::vector<Fp>& secret,
                       const LweParams& lwe, uint64_t last_row_seed);

// Synthetic programs for the benchmark configurations.
// kind: 1 = scalar (PUSH/READ/ADD/MUL), 2 = ciphertext (READ2/READ/SMUL/ADD2/SADD), 3 = mixed.
// The returned program compiles to a length Lp with 2^(log_n-2) <= Lp < 2^(log_n-1), i.e. trace length 2^log_n.
struct SyntheticCase {
    Program program;
    std::vector<uint8_t> pub;
    std::vector<Fp> secret;
};
SyntheticCase make_synthetic(int kind, unsigned log_n, const LweParams& lwe, uint64_t seed);

// LWE client side (fhe/src/server_key.rs:19-76) with a deterministic PRNG instead of thread_rng.
struct LweKey {
    std::vector<Fp> key;
    LweParams params;
};
LweKey lwe_keygen(const LweParams& p, uint64_t seed);
std::vector<Fp> lwe_encrypt(const LweKey& k, uint8_t value, uint64_t seed);
uint8_t lwe_decrypt(const LweKey& k, const Fp* ciphertext);

// The two host field products side by side, for the self-test of the C ABI (ezk_selftest_host_field): the portable
// operator* of f128_host.h and the x86-64 block the Rescue sponge uses (identical by construction; checked against
// big integers in tests/test_host_cpu.py).
void host_field_products(Fp a, Fp b, Fp& portable, Fp& sponge, Fp& squared, Fp& inv_alpha_power);

struct SplitMix64 {
    uint64_t s;
    explicit SplitMix64(uint64_t seed) : s(seed) {}
    uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ULL);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        return z ^ (z >> 31);
    }
    Fp next_fp() {  // uniform in [0, M) by rejection
        for (;;) {
            u128 v = ((u128)next() << 64) | next();
            if (v < Fp::modulus()) return Fp(v);
        }
    }
    Fp next_fp_nonzero() {
        for (;;) {
            Fp v = next_fp();
            if (!v.is_zero()) return v;
        }
    }
};

}  // namespace ezk
