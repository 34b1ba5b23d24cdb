// Is the FP64 pipe a resource of its own next to the integer pipes?
//
// in the integer-bound kernels the ALU pipe and the FMA-heavy pipe (IMAD.WIDE) behave as one
// conserved resource (time tracks the executed integer instruction count, ~0.6 warp-instructions per cycle per
// sub-partition).  B200 keeps a full-rate FP64 pipe (DFMA: 16 lanes per sub-partition, 2 cycles per warp
// instruction) that these kernels leave idle.  A 26-bit-limb product of two field elements is exact in doubles
// (26 x 26 bits = 52 bits per partial product), so IF DFMA issues beside IMAD.WIDE / IADD3 without slowing them,
// part of the butterflies of an NTT pass could run on the FP64 pipe.  This benchmark answers the "if": it times
// DFMA alone, the integer instructions alone, and mixes of them in the same thread (8 independent accumulators
// per kind, 8 warps per sub-partition), in cycles per warp-instruction per sub-partition.  A mix that costs about
// max(parts) means independent pipes; about sum(parts) means one shared issue resource.
//
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_pipe_ubench fp64_pipe_ubench.cu && ./fp64_pipe_ubench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

// MIX bits: 1 = DFMA, 2 = IMAD.WIDE.U32 carry-out + IMAD.WIDE.U32.X carry-in (2 instructions), 4 = IADD3, 8 = DADD pair (the magic-number split hi = (x + C) - C)
template <int MIX>
__global__ void __launch_bounds__(256) kern(uint32_t* out, int iters, uint32_t a, uint32_t b, double fa, double fb) {
    uint32_t lo[8], hi[8], s[8];
    double d[8], e[8];
    for (int c = 0; c < 8; c++) {
        lo[c] = threadIdx.x + c, hi[c] = blockIdx.x * 7 + c;
        s[c] = threadIdx.x * 3 + c;
        d[c] = (double)(threadIdx.x + c) * 1e-3;
        e[c] = (double)(blockIdx.x + c);
    }
    const uint32_t x = a + threadIdx.x, y = b | 1;
    const double fx = fa + threadIdx.x * 1e-9, fy = fb, magic = 6755399441055744.0;  // 1.5 * 2^52
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int c = 0; c < 8; c++) {
            if (MIX & 1) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(d[c]) : "d"(fx), "d"(fy));
            // the integer load is the pattern of the field product: IMAD.WIDE.U32 with carry-out followed by
            // IMAD.WIDE.U32.X with carry-in (ptxas keeps these; a plain `mad.wide.u32` accumulate with loop-invariant
            // multiplicands is hoisted, and with a data-dependent one it is split into IMAD.WIDE + IADD3 + IADD3.X)
            if (MIX & 2)
                asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;\n\t"
                             "madc.lo.cc.u32 %0, %3, %3, %0;\n\tmadc.hi.u32 %1, %3, %3, %1;"
                             : "+r"(lo[c]), "+r"(hi[c]) : "r"(x), "r"(y));
            if (MIX & 4) asm volatile("add.u32 %0, %0, %1;\n\tadd.u32 %0, %0, %2;" : "+r"(s[c]) : "r"(x), "r"(y));
            if (MIX & 8) asm volatile("add.rn.f64 %0, %0, %1;\n\tsub.rn.f64 %0, %0, %1;" : "+d"(e[c]) : "d"(magic));
        }
    }
    uint32_t acc = 0;
    for (int c = 0; c < 8; c++)
        acc += lo[c] ^ hi[c] ^ s[c] ^ (uint32_t)__double2ll_rn(d[c] * 1e-300) ^ (uint32_t)__double2ll_rn(e[c]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MIX>
double run(const char* name, int instr_per_op) {
    int sms = 0, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const int blocks = sms * 4, iters = 4000;
    uint32_t* out;
    cudaMalloc(&out, (size_t)blocks * 256 * 4);
    kern<MIX><<<blocks, 256>>>(out, 10, 3, 5, 1.0000001, 1e-7);
    cudaEvent_t t0, t1;
    cudaEventCreate(&t0), cudaEventCreate(&t1);
    cudaEventRecord(t0);
    kern<MIX><<<blocks, 256>>>(out, iters, 0x12345, 0x6789b, 1.0000001, 1e-7);
    cudaEventRecord(t1);
    cudaEventSynchronize(t1);
    float ms = 0;
    cudaEventElapsedTime(&ms, t0, t1);
    // per sub-partition: blocks * 8 warps / (sms * 4) warps, each iters * 8 ops
    const double ops_per_smsp = (double)blocks * 8 * iters * 8 / (sms * 4);
    const double cycles_per_op = ms * 1e-3 * (khz * 1e3) / ops_per_smsp;
    printf("%-44s %8.3f ms  %6.2f cycles per op (%d instr) per SMSP = %.2f per instruction (at %.3f GHz nominal)\n", name, ms,
           cycles_per_op, instr_per_op, cycles_per_op / instr_per_op, khz * 1e-6);
    cudaFree(out);
    return cycles_per_op;
}

int main() {
    const double dfma = run<1>("DFMA", 1);
    const double wide = run<2>("IMAD.WIDE.U32 + IMAD.WIDE.U32.X (carry)", 2);
    const double iadd = run<4>("IADD3", 1);
    const double dadd = run<8>("DADD + DADD (magic-number split)", 2);
    const double m12 = run<3>("DFMA + IMAD.WIDE pair", 3);
    const double m14 = run<5>("DFMA + IADD3", 2);
    const double m24 = run<6>("IMAD.WIDE pair + IADD3", 3);
    const double m124 = run<7>("DFMA + IMAD.WIDE pair + IADD3", 4);
    const double m1248 = run<15>("DFMA + IMAD.WIDE pair + IADD3 + 2 DADD", 6);
    printf("\nshared-resource model (sum of parts) vs independent pipes (max of parts) vs measured:\n");
    printf("  DFMA + IMAD.WIDE          sum %.2f  max %.2f  measured %.2f\n", dfma + wide, dfma > wide ? dfma : wide, m12);
    printf("  DFMA + IADD3              sum %.2f  max %.2f  measured %.2f\n", dfma + iadd, dfma > iadd ? dfma : iadd, m14);
    printf("  IMAD.WIDE + IADD3         sum %.2f  max %.2f  measured %.2f\n", wide + iadd, wide > iadd ? wide : iadd, m24);
    printf("  DFMA + IMAD.WIDE + IADD3  sum %.2f  measured %.2f\n", dfma + wide + iadd, m124);
    printf("  all five                  sum %.2f  measured %.2f\n", dfma + wide + iadd + dadd, m1248);
    printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
