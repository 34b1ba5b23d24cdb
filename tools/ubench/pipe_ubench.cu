// Issue cost of the integer instructions the f128 arithmetic is made of (cycles per warp-instruction per SM
// sub-partition), measured with 8 independent accumulators per thread and 8 warps per sub-partition.
//
// CHECK THE SASS BEFORE READING A LINE (cuobjdump -sass pipe_ubench | grep -c <mnemonic> per kernel): ptxas rewrites
// loops whose operands are loop-invariant.  As compiled by nvcc 12.9 the loops of OP 1 (IMAD.WIDE carry pair), 2
// (IMAD), 4 (IADD3), 5 (IADD3 + IADD3.X / IMAD.X chain) and 9 hold the instructions they name (232 per kernel =
// 29 x 8), but OP 0's product is hoisted out of the loop (the loop is left with an add pair), and OP 3, 6, 7 and 10
// are partly folded.  The cost of one IMAD.WIDE.U32 (any carry variant) is therefore the OP 1 figure divided by its
// two instructions; a 64-bit multiply-accumulate without carries cannot be forced from PTX (with a data-dependent
// multiplicand ptxas splits it into IMAD.WIDE + IADD3 + IADD3.X).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int OP>
__global__ void __launch_bounds__(256) kern(uint32_t* out, int iters, uint32_t a, uint32_t b) {
    uint32_t lo[8], hi[8];
    for (int c = 0; c < 8; c++) lo[c] = threadIdx.x + c, hi[c] = blockIdx.x * 7 + c;
    uint32_t x = a + threadIdx.x, y = b | 1;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int c = 0; c < 8; c++) {
            if (OP == 0) {  // IMAD.WIDE.U32 (64-bit accumulate, no carry)
                uint64_t acc = ((uint64_t)hi[c] << 32) | lo[c];
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(x), "r"(y));
                lo[c] = (uint32_t)acc, hi[c] = (uint32_t)(acc >> 32);
            }
            if (OP == 1) {  // IMAD.WIDE.U32 with carry out + .X with carry in (two fused pairs)
                asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;\n\t"
                             "madc.lo.cc.u32 %0, %3, %3, %0;\n\tmadc.hi.u32 %1, %3, %3, %1;"
                             : "+r"(lo[c]), "+r"(hi[c]) : "r"(x), "r"(y));
            }
            if (OP == 2) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(lo[c]) : "r"(x), "r"(y));       // IMAD
            if (OP == 3) asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(lo[c]) : "r"(x), "r"(y));       // IMAD.HI
            if (OP == 4) asm volatile("add.u32 %0, %0, %1;\n\tadd.u32 %0, %0, %2;" : "+r"(lo[c]) : "r"(x), "r"(y));  // IADD3
            if (OP == 5) {  // 4 adds with carry chain (IADD3 + 3 IADD3.X)
                asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.cc.u32 %1, %1, %3;\n\taddc.cc.u32 %0, %0, %3;\n\taddc.u32 %1, %1, %2;"
                             : "+r"(lo[c]), "+r"(hi[c]) : "r"(x), "r"(y));
            }
            if (OP == 6) asm volatile("xor.b32 %0, %0, %1;\n\tand.b32 %0, %0, %2;" : "+r"(lo[c]) : "r"(x), "r"(y));  // LOP3
            if (OP == 8) {  // carry-OUT only: IMAD.WIDE.U32 R, P0 + IADD3.X consuming it
                asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;\n\taddc.u32 %0, %0, 0;"
                             : "+r"(lo[c]), "+r"(hi[c]) : "r"(x), "r"(y));
            }
            if (OP == 9) {  // carry-IN only: IADD3 producing a carry + IMAD.WIDE.U32.X
                asm volatile("add.cc.u32 %0, %0, %3;\n\tmadc.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;"
                             : "+r"(lo[c]), "+r"(hi[c]) : "r"(x), "r"(y));
            }
            if (OP == 10) {  // reference for 8/9: one IADD3 with carry-out + one IADD3.X
                asm volatile("add.cc.u32 %0, %0, %3;\n\taddc.u32 %1, %1, %2;" : "+r"(lo[c]), "+r"(hi[c]) : "r"(x), "r"(y));
            }
            if (OP == 7) {  // 32x32 -> 64 from separate lo / hi multiplies
                asm volatile("mad.lo.u32 %0, %2, %3, %0;\n\tmad.hi.u32 %1, %2, %3, %1;" : "+r"(lo[c]), "+r"(hi[c]) : "r"(x), "r"(y));
            }
        }
    }
    uint32_t acc = 0;
    for (int c = 0; c < 8; c++) acc += lo[c] ^ hi[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int OP>
void run(const char* name, int instr_per_op) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int blocks = sms * 4, iters = 4000;
    uint32_t* out;
    cudaMalloc(&out, (size_t)blocks * 256 * 4);
    kern<OP><<<blocks, 256>>>(out, 10, 3, 5);
    cudaEvent_t a, b;
    cudaEventCreate(&a), cudaEventCreate(&b);
    cudaEventRecord(a);
    kern<OP><<<blocks, 256>>>(out, iters, 0x12345, 0x6789b);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    const double warp_instr_per_smsp = (double)blocks * 8 * iters * 8 * instr_per_op / (sms * 4);
    printf("%-28s %8.3f ms   %.2f cycles per warp-instruction per SMSP (at 1.965 GHz)\n", name, ms, ms * 1e-3 * 1.965e9 / warp_instr_per_smsp);
    cudaFree(out);
}

int main() {
    run<0>("(void: product hoisted, add pair left)", 1);
    run<1>("IMAD.WIDE.U32 carry (x2)", 2);
    run<2>("IMAD (lo)", 1);
    run<3>("IMAD.HI", 1);
    run<7>("IMAD lo + IMAD.HI", 2);
    run<8>("WIDE carry-out + IADD3.X", 2);
    run<9>("IADD3.cc + WIDE.X carry-in", 2);
    run<10>("IADD3.cc + IADD3.X", 2);
    run<4>("IADD3 (2 adds -> 1)", 1);
    run<5>("IADD3 + 3 IADD3.X chain", 4);
    run<6>("LOP3 (xor+and -> 1)", 1);
    printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
