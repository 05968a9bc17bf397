// imad_wide_forms.cu -- does the operand form of the Philox multiply matter?  IMAD.WIDE.U32 with the multiplier as an
// immediate, as a uniform (kernel-parameter) value, or as a per-thread register; also mul.hi + mul.lo and mad.wide.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ITER = 4096, CH = 8;

template <int KIND>
__global__ void __launch_bounds__(256, 5) k(uint32_t *out, uint32_t mult, uint32_t seed)
{
    uint32_t a[CH], b[CH];
    const uint32_t mreg = mult + (threadIdx.x & 1) * 2;      // per-thread multiplier (odd)
#pragma unroll
    for (int i = 0; i < CH; ++i) { a[i] = seed + threadIdx.x * 7 + i; b[i] = seed ^ (i * 0x9E3779B9u); }
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            unsigned long long p;
            if (KIND == 0) p = (unsigned long long)a[i] * 0xD2511F53u;
            else if (KIND == 1) p = (unsigned long long)a[i] * mult;
            else if (KIND == 2) p = (unsigned long long)a[i] * mreg;
            else if (KIND == 3) p = (unsigned long long)a[i] * 0xD2511F53u + b[i];          // mad.wide
            else { p = ((unsigned long long)__umulhi(a[i], 0xD2511F53u) << 32) | (uint32_t)(a[i] * 0xD2511F53u); }
            a[i] = (uint32_t)(p >> 32) ^ b[i];
            b[i] = (uint32_t)p;
        }
    }
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) x ^= a[i] ^ b[i];
    if (x == 0x12345) out[0] = x;
}

template <int KIND> void run(const char *name, uint32_t *d, int sms, double ghz)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = sms * 5;
    k<KIND><<<blocks, 256>>>(d, 0xD2511F53u, 1); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<KIND><<<blocks, 256>>>(d, 0xD2511F53u, 2); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double pairs_per_smsp = (double)blocks * 8.0 * ITER * CH / (sms * 4.0);
    printf("%-46s %7.3f ms  %5.2f cycles per (multiply + xor) per SMSP\n", name, ms, ms * 1e-3 * ghz * 1e9 / pairs_per_smsp);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    uint32_t *d; cudaMalloc(&d, 4);
    run<0>("mul.wide.u32, immediate multiplier", d, p.multiProcessorCount, khz * 1e-6);
    run<1>("mul.wide.u32, uniform (parameter) multiplier", d, p.multiProcessorCount, khz * 1e-6);
    run<2>("mul.wide.u32, per-thread register multiplier", d, p.multiProcessorCount, khz * 1e-6);
    run<3>("mad.wide.u32, immediate multiplier", d, p.multiProcessorCount, khz * 1e-6);
    run<4>("mul.hi.u32 + mul.lo.u32", d, p.multiProcessorCount, khz * 1e-6);
    return 0;
}
