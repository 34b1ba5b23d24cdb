// One tiny kernel per field primitive, for tools/sass_primitives.py: the SASS between the loads and the store of each
// is what DESIGN.md section 4 counts (instructions per product / addition / butterfly).  Not part of the library.
#include "../../encrypt_zkvm_b200/csrc/field/f128.cuh"
using namespace ezk::dev;

#define PROBE(NAME, BODY)                                                                              \
    extern "C" __global__ void NAME(const uint4* a, const uint4* b, uint4* out, uint32_t* flag) {       \
        const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;                                       \
        fe x = fe_load(a + t), y = fe_load(b + t);                                                      \
        uint32_t rare = 0;                                                                              \
        fe r;                                                                                           \
        BODY;                                                                                           \
        fe_store(out + t, r);                                                                           \
        if (rare == 0xFFFFFFFFu) *flag = 1;                                                             \
    }

PROBE(probe_mul_flag_carry, r = fe_mul_flag<0>(x, y, rare))        // strided NTT pass
PROBE(probe_mul_flag_lean, r = fe_mul_flag<1>(x, y, rare))         // final NTT pass, constraint kernel
PROBE(probe_mul_exact, r = fe_mul(x, y))                           // exact product (rare redo paths, small kernels)
PROBE(probe_add_flag_masked, r = fe_add_flag_masked(x, y, rare))   // strided NTT pass
PROBE(probe_add_flag_predicated, r = fe_add_flag(x, y, rare))      // final NTT pass, constraint kernel
PROBE(probe_sub, r = fe_sub(x, y))

// precomputed-form product: the table operand is 4 x 16 bytes
extern "C" __global__ void probe_mul_pre_flag(const uint4* a, const uint4* w, uint4* out, uint32_t* flag) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    fe x = fe_load(a + t);
    fe_pre W = fe_pre_load(w + 4 * t);
    uint32_t rare = 0;
    fe r = fe_mul_pre_flag<1>(x, W, rare);
    fe_store(out + t, r);
    if (rare == 0xFFFFFFFFu) *flag = 1;
}

// one butterfly with a twiddle from memory: (a, b) -> (a + b, (a - b) * w)
extern "C" __global__ void probe_butterfly(const uint4* a, const uint4* b, const uint4* w, uint4* out, uint32_t* flag) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    fe x = fe_load(a + t), y = fe_load(b + t), tw = fe_load(w + t);
    uint32_t rare = 0;
    fe s = fe_add_flag_masked(x, y, rare);
    fe d = fe_mul_flag<0>(fe_sub(x, y), tw, rare);
    fe_store(out + 2 * t, s);
    fe_store(out + 2 * t + 1, d);
    if (rare == 0xFFFFFFFFu) *flag = 1;
}
