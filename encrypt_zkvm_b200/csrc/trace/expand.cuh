// Device-side generation of the trace columns that are pure functions of the executed operation list (SURVEY 8f-2,
// scoped to what parallelises).  Of the 28 columns `Processor::trace` assembles (vm/src/processor/mod.rs:76-84):
//   column 0      clk            = row index                                   (system.rs:19-35)
//   columns 1..5  op bits        = bit k of the op executed at that row, 0 after the program   (decoder.rs:31-76)
//   column 6      chiplet flag   = 1 while program operations run, 0 after      (chiplets.rs:99-105)
//   column 11     stack depth    = prefix sum of the per-operation depth change, last value repeated (stack.rs:80-97,278-280)
// need only the operation codes (one byte per executed operation instead of 8 x 16 bytes per row); the Rescue sponge
// columns (one 128-bit exponentiation per lane per operation, each depending on the previous state) and the stack
// registers (a sequential machine over field values) stay with the host VM, as the north star says.
// Row n-1 of EVERY column is overwritten with caller-supplied values (the reference's thread_rng row, mod.rs:86-92).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace ezk {

constexpr uint32_t kOpColumns[8] = {0, 1, 2, 3, 4, 5, 6, 11};
inline bool is_op_column(uint32_t c) { return c <= 6 || c == 11; }

// d_trace: 28 columns of n elements (pitch n).  d_codes: `count` operation codes (count < n).  d_last_row: 28 elements.
// d_scan: (n / 1024 + 2) uint32 of scratch.  lwe_size = k + 1 (depth change of READ2 / ADD2).
// *d_flag |= 4 when an operation code is not one of the nine opcodes.
int expand_op_columns(cudaStream_t s, const uint8_t* d_codes, uint64_t count, uint64_t n, uint32_t lwe_size,
                      const uint4* d_last_row, uint32_t* d_scan, uint4* d_trace, uint32_t* d_flag);

}  // namespace ezk
