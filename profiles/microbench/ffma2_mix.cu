// ffma2_mix.cu -- does Blackwell's packed FP32 (fma.rn.f32x2 -> FFMA2) relieve the FE kernel's instruction mix?
// Same independent-chain construction as fe_mix_bound.cu; the FP32 part is issued either as NFP scalar FFMAs or as
// NFP/2 FFMA2s (same flops).  Prints cycles per warp-step per SM sub-partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_mix ffma2_mix.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ITER = 4096;

template <int NMUL, int NLOP, int NFP, int NMUFU, bool PACKED>
__global__ void __launch_bounds__(256, 5) mix(uint32_t *out, uint32_t seed)
{
    uint32_t a[8], b[8];
    float f[8], g[4];
    unsigned long long pf[4], pg[2];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = seed + threadIdx.x * 7 + i; b[i] = seed ^ (i * 0x9E3779B9u); f[i] = 1.0f + (threadIdx.x + i) * 1e-3f; }
#pragma unroll
    for (int i = 0; i < 4; ++i) { g[i] = 1.5f + i; pf[i] = ((unsigned long long)__float_as_uint(f[2 * i]) << 32) | __float_as_uint(f[2 * i + 1]); }
    pg[0] = ((unsigned long long)__float_as_uint(0.999f) << 32) | __float_as_uint(0.998f);
    pg[1] = ((unsigned long long)__float_as_uint(1e-3f) << 32) | __float_as_uint(2e-3f);
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < NMUL; ++i)
            asm volatile("{ .reg .u64 t; mul.wide.u32 t, %0, 0xD2511F53; mov.b64 {%0, %1}, t; }" : "+r"(a[i % 8]), "+r"(b[i % 8]));
#pragma unroll
        for (int i = 0; i < NLOP; ++i) asm volatile("lop3.b32 %0, %0, %1, 0x12345678, 0x96;" : "+r"(a[(i + 3) % 8]) : "r"(b[(i + 5) % 8]));
#pragma unroll
        for (int i = 0; i < 2; ++i) asm volatile("shr.u32 %0, %0, 1;" : "+r"(b[(i + 1) % 8]));
        if (PACKED) {
#pragma unroll
            for (int i = 0; i < NFP / 2; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(pf[i % 4]) : "l"(pg[0]), "l"(pg[1]));
        } else {
#pragma unroll
            for (int i = 0; i < NFP; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(f[i % 8]) : "f"(g[i % 4]));
        }
#pragma unroll
        for (int i = 0; i < NMUFU; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(g[i % 4]));
    }
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) x ^= a[i] ^ b[i] ^ __float_as_uint(f[i]);
#pragma unroll
    for (int i = 0; i < 4; ++i) x ^= (uint32_t)pf[i] ^ (uint32_t)(pf[i] >> 32);
    x ^= __float_as_uint(g[0] + g[1] + g[2] + g[3]);
    if (x == 0x12345) out[0] = x;
}

template <int NMUL, int NLOP, int NFP, int NMUFU, bool PACKED> void run(const char *name, uint32_t *d, int sms, double ghz)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = sms * 5;
    mix<NMUL, NLOP, NFP, NMUFU, PACKED><<<blocks, 256>>>(d, 1); cudaDeviceSynchronize();
    cudaEventRecord(e0); mix<NMUL, NLOP, NFP, NMUFU, PACKED><<<blocks, 256>>>(d, 2); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double warp_steps_per_smsp = (double)blocks * 8.0 * ITER / (sms * 4.0);
    const double cyc = ms * 1e-3 * ghz * 1e9 / warp_steps_per_smsp;
    printf("%-64s %7.3f ms  %6.2f cycles per warp-step per SMSP\n", name, ms, cyc);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    uint32_t *d; cudaMalloc(&d, 4);
    const int sms = p.multiProcessorCount; const double ghz = khz * 1e-6;
    run<0, 0, 12, 0, false>("12 FFMA only", d, sms, ghz);
    run<0, 0, 12, 0, true>("6 FFMA2 only (same flops)", d, sms, ghz);
    run<0, 0, 24, 0, true>("12 FFMA2 only", d, sms, ghz);
    run<8, 9, 12, 4, false>("FE mix: 8 IMAD.WIDE, 9 LOP3, 2 SHF, 4 MUFU + 12 FFMA", d, sms, ghz);
    run<8, 9, 12, 4, true>("FE mix: ... + 6 FFMA2", d, sms, ghz);
    run<8, 9, 0, 4, false>("FE mix without FP32", d, sms, ghz);
    run<8, 0, 12, 0, false>("8 IMAD.WIDE + 12 FFMA", d, sms, ghz);
    run<8, 0, 12, 0, true>("8 IMAD.WIDE + 6 FFMA2", d, sms, ghz);
    run<6, 10, 12, 4, false>("dense FE mix: 6 IMAD.WIDE, 10 LOP3, 4 MUFU + 12 FFMA", d, sms, ghz);
    run<6, 10, 12, 4, true>("dense FE mix: ... + 6 FFMA2", d, sms, ghz);
    run<14, 18, 24, 9, false>("EM trial mix: 14 IMAD.WIDE, 18 LOP3, 9 MUFU + 24 FFMA", d, sms, ghz);
    run<14, 18, 24, 9, true>("EM trial mix: ... + 12 FFMA2", d, sms, ghz);
    return 0;
}
