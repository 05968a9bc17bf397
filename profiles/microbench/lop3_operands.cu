// lop3_operands.cu -- does the operand kind of the Philox key XOR matter?  LOP3 with an immediate, a
// uniform-register (kernel parameter) or a per-thread register third operand; and IMAD.WIDE + LOP3 mixes.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ITER = 2048, CH = 8;
struct Keys { uint32_t k[16]; };

template <int KIND>
__global__ void __launch_bounds__(1024) k(uint32_t *out, const __grid_constant__ Keys K, uint32_t seed)
{
    uint32_t a[CH], b[CH], r[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) { a[i] = seed + threadIdx.x * 7 + i; b[i] = seed ^ (i * 0x9E3779B9u); r[i] = a[i] * 3 + 1; }
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            if (KIND == 0) {            // immediate key
                a[i] = a[i] ^ b[i] ^ 0x12345678u; b[i] = b[i] ^ a[i] ^ 0x9abcdef1u;
            } else if (KIND == 1) {     // uniform key from the parameter block
                a[i] = a[i] ^ b[i] ^ K.k[i]; b[i] = b[i] ^ a[i] ^ K.k[i + 8];
            } else if (KIND == 2) {     // per-thread register key
                a[i] = a[i] ^ b[i] ^ r[i]; b[i] = b[i] ^ a[i] ^ r[(i + 1) % CH];
            } else if (KIND == 3) {     // Philox-like: wide multiply + xor with immediate keys
                unsigned long long p = (unsigned long long)0xD2511F53u * a[i];
                a[i] = (uint32_t)(p >> 32) ^ b[i] ^ 0x12345678u; b[i] = (uint32_t)p;
            } else if (KIND == 4) {     // Philox-like: wide multiply + xor with uniform keys
                unsigned long long p = (unsigned long long)0xD2511F53u * a[i];
                a[i] = (uint32_t)(p >> 32) ^ b[i] ^ K.k[i]; b[i] = (uint32_t)p;
            } else if (KIND == 5) {     // Philox-like: wide multiply + xor with register keys
                unsigned long long p = (unsigned long long)0xD2511F53u * a[i];
                a[i] = (uint32_t)(p >> 32) ^ b[i] ^ r[i]; b[i] = (uint32_t)p;
            }
        }
    }
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) x ^= a[i] ^ b[i];
    if (x == 0x12345) out[0] = x;
}

template <int KIND> void run(const char *name, int ipi, uint32_t *d, const Keys &K, int sms, double ghz)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = sms * 2;
    k<KIND><<<blocks, 1024>>>(d, K, 1); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<KIND><<<blocks, 1024>>>(d, K, 2); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double wi = (double)blocks * 32.0 * ITER * CH * ipi;
    printf("%-44s %8.3f ms  %6.3f warp-instr/clk/SM\n", name, ms, wi / (ms * 1e-3 * ghz * 1e9) / sms);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    Keys K; for (int i = 0; i < 16; ++i) K.k[i] = 0x9E3779B9u * (i + 1);
    uint32_t *d; cudaMalloc(&d, 4);
    run<0>("LOP3 x2, immediate key", 2, d, K, p.multiProcessorCount, khz * 1e-6);
    run<1>("LOP3 x2, uniform (param) key", 2, d, K, p.multiProcessorCount, khz * 1e-6);
    run<2>("LOP3 x2, register key", 2, d, K, p.multiProcessorCount, khz * 1e-6);
    run<3>("IMAD.WIDE + LOP3, immediate key", 2, d, K, p.multiProcessorCount, khz * 1e-6);
    run<4>("IMAD.WIDE + LOP3, uniform key", 2, d, K, p.multiProcessorCount, khz * 1e-6);
    run<5>("IMAD.WIDE + LOP3, register key", 2, d, K, p.multiProcessorCount, khz * 1e-6);
    return 0;
}
