// FRI fold-and-commit (kernel K9 of SURVEY 8a').  Replaces winter-fri's FriProver::build_layer<8> /
// apply_drp / set_remainder, reached from `Prover::prove` (vm/src/lib.rs:26).
// Folding factor 8; the domain offset stays 3 at every layer (Winterfell's behaviour).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../common.h"

namespace ezk {

struct FriFoldConsts {
    uint64_t zinv[4][2];    // zeta^-k, k = 0..3, zeta = w_8
    uint64_t inv8[2];       // 1/8
    uint64_t alpha_oinv[2]; // alpha / 3
};

// next[i] = sum_k (alpha / x_i)^k * (1/8) sum_j evals[i + j*m] zeta^(-jk),  x_i = 3 w_s^i,  i < m = s/8
// multi-GPU (sh): evals / next hold the positions this rank owns (p mod world = rank) in ascending order; log_s is the
// size of the whole layer
int fri_fold(cudaStream_t s, const uint4* root_inv, const uint4* evals, uint32_t log_s, FriFoldConsts c, uint4* next,
             RowShard sh = RowShard());

// coeffs[k] = (1/s) 3^-k sum_i evals[i] w_s^(-ik), k < s (s <= 4096): remainder interpolation over the coset
int fri_remainder(cudaStream_t s, const uint4* root_inv, const uint4* off_inv, const uint4* evals, uint32_t log_s,
                  const uint64_t inv_s[2], uint4* coeffs);

}  // namespace ezk
