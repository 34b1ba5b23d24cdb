// TEST INFRASTRUCTURE (part of oracle/liborc.so; see oracle/capi.cpp): input generator of the CPU legs.
//
// `bench.py --impl reference` proves the benchmark's synthetic program with the CPU oracle and must not load the
// product library to get its input, so the host VM harness (encrypt_zkvm_b200/csrc/host/vm.cc - the restatement of
// the reference's `vm` crate: Program::compile vm/src/program/mod.rs:37-96, Processor::trace
// vm/src/processor/mod.rs:61-95) is compiled into the oracle library a second time, under its own entry point.
// Nothing here is on the measured path: the trace is built before the timed region starts.
#include "../encrypt_zkvm_b200/csrc/host/vm.cc"
#include <cstring>

extern "C" {

// Builds BASELINE.md's synthetic case (kind 1 scalar / 2 ciphertext / 3 mixed, trace length 2^log_n) with the same
// generator and seeds as ezk_synthetic_case.  trace_out: 28 columns of 2^log_n elements (16 little-endian bytes each,
// column-major); pub_out: program_hash[2] ++ stack_outputs[16].  Returns 0, or -1 with nothing written on a VM error.
int orc_synthetic_trace(int kind, unsigned log_n, unsigned lwe_k, unsigned lwe_delta, unsigned long long seed,
                        unsigned char* trace_out, unsigned char* pub_out) {
    try {
        ezk::LweParams lwe;
        lwe.k = lwe_k, lwe.delta = lwe_delta;
        ezk::SyntheticCase c = ezk::make_synthetic(kind, log_n, lwe, seed);
        ezk::ExecutionTrace t = ezk::execute(c.program, c.pub, c.secret, lwe, seed ^ 0x5EEDULL);
        if (t.n != ((size_t)1 << log_n) || t.columns.size() != 28) return -1;
        for (size_t col = 0; col < 28; col++) memcpy(trace_out + col * t.n * 16, t.columns[col].data(), t.n * 16);
        ezk::fp_store(pub_out, c.program.hash[0]);
        ezk::fp_store(pub_out + 16, c.program.hash[1]);
        for (int i = 0; i < 16; i++) ezk::fp_store(pub_out + 32 + 16 * i, t.outputs[i]);
        return 0;
    } catch (...) {
        return -1;
    }
}

}  // extern "C"
