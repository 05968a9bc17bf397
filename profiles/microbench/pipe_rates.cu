// pipe_rates.cu -- issue-rate microbenchmarks behind the FE kernel's instruction budget.
// Measures warp-instructions per clock per SM for the instruction kinds in the Philox / Box-Muller loop.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rates pipe_rates.cu && ./pipe_rates
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITER = 2048;
constexpr int CHAINS = 8;

template <int KIND>
__global__ void __launch_bounds__(1024) k(uint32_t *out, uint32_t seed)
{
    uint32_t a[CHAINS], b[CHAINS];
    float f[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) {
        a[i] = seed + threadIdx.x * 7 + i;
        b[i] = seed ^ (i * 0x9E3779B9u);
        f[i] = 1.0f + (float)(threadIdx.x + i) * 1e-3f;
    }
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) {
            if (KIND == 0) {          // IMAD.WIDE.U32
                asm volatile("{ .reg .u64 t; mul.wide.u32 t, %0, 0xD2511F53; mov.b64 {%0, %1}, t; }" : "+r"(a[i]), "+r"(b[i]));
            } else if (KIND == 1) {   // IMAD.HI.U32
                asm volatile("mul.hi.u32 %0, %0, 0xD2511F53;" : "+r"(a[i]));
            } else if (KIND == 2) {   // IMAD (low 32)
                asm volatile("mul.lo.u32 %0, %0, 0xD2511F53;" : "+r"(a[i]));
            } else if (KIND == 3) {   // LOP3
                asm volatile("lop3.b32 %0, %0, %1, 0x12345678, 0x96;" : "+r"(a[i]) : "r"(b[i]));
            } else if (KIND == 4) {   // IMAD.WIDE + LOP3 (one Philox half-round)
                asm volatile("{ .reg .u64 t; mul.wide.u32 t, %0, 0xD2511F53; mov.b64 {%0, %1}, t; }" : "+r"(a[i]), "+r"(b[i]));
                asm volatile("lop3.b32 %0, %0, %1, 0x12345678, 0x96;" : "+r"(a[i]) : "r"(b[i]));
            } else if (KIND == 5) {   // FFMA
                asm volatile("fma.rn.f32 %0, %0, 0f3F7FFFF0, 0f3A000000;" : "+f"(f[i]));
            } else if (KIND == 6) {   // MUFU.EX2
                asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
            } else if (KIND == 7) {   // IMAD.WIDE + FFMA
                asm volatile("{ .reg .u64 t; mul.wide.u32 t, %0, 0xD2511F53; mov.b64 {%0, %1}, t; }" : "+r"(a[i]), "+r"(b[i]));
                asm volatile("fma.rn.f32 %0, %0, 0f3F7FFFF0, 0f3A000000;" : "+f"(f[i]));
            } else if (KIND == 8) {   // mul.hi + mul.lo + LOP3 (half-round with split multiply)
                uint32_t hi;
                asm volatile("mul.hi.u32 %0, %1, 0xD2511F53;" : "=r"(hi) : "r"(a[i]));
                asm volatile("mul.lo.u32 %0, %0, 0xD2511F53;" : "+r"(a[i]));
                asm volatile("lop3.b32 %0, %1, %2, 0x12345678, 0x96;" : "=r"(b[i]) : "r"(hi), "r"(b[i]));
            } else if (KIND == 9) {   // IMAD.WIDE + LOP3 + FFMA + FFMA (the kernel's rough mix)
                asm volatile("{ .reg .u64 t; mul.wide.u32 t, %0, 0xD2511F53; mov.b64 {%0, %1}, t; }" : "+r"(a[i]), "+r"(b[i]));
                asm volatile("lop3.b32 %0, %0, %1, 0x12345678, 0x96;" : "+r"(a[i]) : "r"(b[i]));
                asm volatile("fma.rn.f32 %0, %0, 0f3F7FFFF0, 0f3A000000;" : "+f"(f[i]));
            } else if (KIND == 10) {  // MUFU + 8 FFMA (XU overlap)
                asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
                asm volatile("fma.rn.f32 %0, %0, 0f3F7FFFF0, 0f3A000000;" : "+f"(f[i]));
                asm volatile("fma.rn.f32 %0, %0, 0f3F7FFFF0, 0f3A000000;" : "+f"(f[i]));
                asm volatile("fma.rn.f32 %0, %0, 0f3F7FFFF0, 0f3A000000;" : "+f"(f[i]));
            }
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) r ^= a[i] ^ b[i] ^ __float_as_uint(f[i]);
    if (r == 0x12345) out[0] = r;
}

template <int KIND>
void run(const char *name, int instr_per_chain_iter, uint32_t *d, int sms, double ghz)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int blocks = sms * 2;
    k<KIND><<<blocks, 1024>>>(d, 1);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<KIND><<<blocks, 1024>>>(d, 2);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double warp_instr = (double)blocks * 32.0 * ITER * CHAINS * instr_per_chain_iter;
    const double per_clk_sm = warp_instr / (ms * 1e-3 * ghz * 1e9) / sms;
    printf("%-44s %8.3f ms  %6.3f warp-instr/clk/SM  (%5.2f thread-instr/clk/SM)\n", name, ms, per_clk_sm, per_clk_sm * 32);
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double ghz = khz * 1e-6;
    printf("%s, %d SMs, %.3f GHz nominal\n", p.name, p.multiProcessorCount, ghz);
    uint32_t *d;
    cudaMalloc(&d, 4);
    const int sms = p.multiProcessorCount;
    run<0>("IMAD.WIDE.U32", 1, d, sms, ghz);
    run<1>("IMAD.HI.U32", 1, d, sms, ghz);
    run<2>("IMAD (mul.lo)", 1, d, sms, ghz);
    run<3>("LOP3", 1, d, sms, ghz);
    run<4>("IMAD.WIDE + LOP3", 2, d, sms, ghz);
    run<5>("FFMA", 1, d, sms, ghz);
    run<6>("MUFU.EX2", 1, d, sms, ghz);
    run<7>("IMAD.WIDE + FFMA", 2, d, sms, ghz);
    run<8>("IMAD.HI + IMAD.LO + LOP3", 3, d, sms, ghz);
    run<9>("IMAD.WIDE + LOP3 + FFMA", 3, d, sms, ghz);
    run<10>("MUFU + 3 FFMA", 4, d, sms, ghz);
    return 0;
}
