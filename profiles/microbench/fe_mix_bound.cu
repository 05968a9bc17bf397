// fe_mix_bound.cu -- upper bound for the FE kernel's instruction MIX on one SM sub-partition: the same counts per
// path-step (8 IMAD.WIDE.U32, 9 LOP3, 2 LEA.HI-like shifts, 12 FP32, 4 MUFU) issued from independent chains, i.e.
// without the Philox / Box-Muller dependency structure.  Prints cycles per warp-step per SMSP.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ITER = 4096;

template <int NMUL, int NLOP, int NFP, int NMUFU>
__global__ void __launch_bounds__(256, 5) mix(uint32_t *out, uint32_t seed)
{
    uint32_t a[8], b[8];
    float f[8], g[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = seed + threadIdx.x * 7 + i; b[i] = seed ^ (i * 0x9E3779B9u); f[i] = 1.0f + (threadIdx.x + i) * 1e-3f; }
#pragma unroll
    for (int i = 0; i < 4; ++i) g[i] = 1.5f + i;
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < NMUL; ++i)
            asm volatile("{ .reg .u64 t; mul.wide.u32 t, %0, 0xD2511F53; mov.b64 {%0, %1}, t; }" : "+r"(a[i % 8]), "+r"(b[i % 8]));
#pragma unroll
        for (int i = 0; i < NLOP; ++i) asm volatile("lop3.b32 %0, %0, %1, 0x12345678, 0x96;" : "+r"(a[(i + 3) % 8]) : "r"(b[(i + 5) % 8]));
#pragma unroll
        for (int i = 0; i < 2; ++i) asm volatile("shr.u32 %0, %0, 1;" : "+r"(b[(i + 1) % 8]));
#pragma unroll
        for (int i = 0; i < NFP; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f[i % 8]) : "f"(g[i % 4]));
#pragma unroll
        for (int i = 0; i < NMUFU; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(g[i % 4]));
    }
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) x ^= a[i] ^ b[i] ^ __float_as_uint(f[i]);
    x ^= __float_as_uint(g[0] + g[1] + g[2] + g[3]);
    if (x == 0x12345) out[0] = x;
}

template <int NMUL, int NLOP, int NFP, int NMUFU> void run(const char *name, uint32_t *d, int sms, double ghz)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = sms * 5;
    mix<NMUL, NLOP, NFP, NMUFU><<<blocks, 256>>>(d, 1); cudaDeviceSynchronize();
    cudaEventRecord(e0); mix<NMUL, NLOP, NFP, NMUFU><<<blocks, 256>>>(d, 2); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double warp_steps_per_smsp = (double)blocks * 8.0 * ITER / (sms * 4.0);
    const double cyc = ms * 1e-3 * ghz * 1e9 / warp_steps_per_smsp;
    printf("%-52s %7.3f ms  %6.2f cycles per warp-step per SMSP  (%d instr)\n", name, ms, cyc, NMUL + NLOP + 2 + NFP + NMUFU);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    uint32_t *d; cudaMalloc(&d, 4);
    const int sms = p.multiProcessorCount; const double ghz = khz * 1e-6;
    run<8, 9, 12, 4>("FE mix: 8 IMAD.WIDE, 9 LOP3, 2 SHF, 12 FP32, 4 MUFU", d, sms, ghz);
    run<8, 9, 12, 0>("... without the MUFU", d, sms, ghz);
    run<8, 9, 0, 4>("... without the FP32", d, sms, ghz);
    run<8, 0, 12, 4>("... without the LOP3", d, sms, ghz);
    run<0, 9, 12, 4>("... without the IMAD.WIDE", d, sms, ghz);
    run<8, 0, 0, 0>("8 IMAD.WIDE only", d, sms, ghz);
    run<7, 9, 12, 4>("dense-like: 6.6 -> 7 IMAD.WIDE", d, sms, ghz);
    // which second limit blends in below ~45 cycles (dense FE mode, EM trial)?  Single pipes and pairs:
    run<0, 9, 12, 0>("9 LOP3 + 2 SHF + 12 FP32", d, sms, ghz);
    run<0, 9, 0, 4>("9 LOP3 + 2 SHF + 4 MUFU", d, sms, ghz);
    run<0, 0, 12, 4>("2 SHF + 12 FP32 + 4 MUFU", d, sms, ghz);
    run<0, 0, 0, 4>("2 SHF + 4 MUFU", d, sms, ghz);
    run<0, 9, 0, 0>("9 LOP3 + 2 SHF", d, sms, ghz);
    run<0, 0, 12, 0>("2 SHF + 12 FP32", d, sms, ghz);
    run<6, 10, 11, 4>("dense FE (3 steps/block): 5.5 -> 6 IMAD.WIDE, 10 LOP3, 11 FP32, 4 MUFU", d, sms, ghz);
    run<5, 10, 11, 4>("... with 5 IMAD.WIDE", d, sms, ghz);
    run<4, 10, 11, 4>("... with 4 IMAD.WIDE", d, sms, ghz);
    run<6, 10, 11, 3>("... 6 IMAD.WIDE, 3 MUFU", d, sms, ghz);
    run<6, 10, 11, 2>("... 6 IMAD.WIDE, 2 MUFU", d, sms, ghz);
    run<14, 18, 24, 9>("EM trial (4 per 3 blocks): 13.5 -> 14 IMAD.WIDE, 18 LOP3, 24 FP32, 9 MUFU", d, sms, ghz);
    run<12, 18, 24, 9>("... with 12 IMAD.WIDE", d, sms, ghz);
    run<18, 18, 24, 9>("... with 18 IMAD.WIDE (one block per trial)", d, sms, ghz);
    return 0;
}
