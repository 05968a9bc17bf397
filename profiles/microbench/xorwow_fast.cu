// xorwow_fast.cu -- feasibility: the native fast-math FE step fed by the XORWOW integer stream (no Philox multiplies),
// P paths per thread, state in registers.  Not product code.  Prints ms for 2^24 paths x 1000 steps.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I ../../nmch_b200/csrc -o xorwow_fast xorwow_fast.cu
#include <cstdio>
#include <cuda_runtime.h>

#include "device_common.cuh"
using namespace nmchb;

struct Consts { float crdt, zr, zc, va, vb, vs, S0, v0, K; int N; };
struct Xw {
    uint32_t d, v0, v1, v2, v3, v4;
    __device__ __forceinline__ uint32_t next()
    {
        const uint32_t t = v0 ^ (v0 >> 2);
        v0 = v1; v1 = v2; v2 = v3; v3 = v4;
        v4 = (v4 ^ (v4 << 4)) ^ (t ^ (t << 1));
        d += 362437u;
        return v4 + d;
    }
};

template <int P, int THREADS, int MINB, int UNROLL>
__global__ void __launch_bounds__(THREADS, MINB) fe(const __grid_constant__ Consts c, double *out)
{
    const uint32_t path0 = blockIdx.x * (P * THREADS) + threadIdx.x;
    float S[P], V[P];
    Xw x[P];
#pragma unroll
    for (int j = 0; j < P; ++j) {
        S[j] = c.S0; V[j] = c.v0;
        const uint32_t p = path0 + j * THREADS;
        x[j] = Xw{p * 2654435761u, p ^ 123456789u, p + 362436069u, p * 3u + 521288629u, p ^ 88675123u, p + 5783321u};
    }
#pragma unroll UNROLL
    for (int n = 0; n < c.N; ++n) {
#pragma unroll
        for (int j = 0; j < P; ++j) {
            const uint32_t wa = x[j].next(), wb = x[j].next();
            const float f1 = bits_to_1_2(wa), f2 = bits_to_1_2(wb);
            const float u = f1 - 0.99999994f;
            const float l2 = lg2_approx(u);
            const float q = sqrt_approx(-(V[j] * l2));
            const float ang = f2 * 6.2831855f;
            const float gs = q * sin_approx(ang), gc = q * cos_approx(ang);
            float m = fmaf(gs, c.zr, c.crdt);
            m = fmaf(gc, c.zc, m);
            S[j] *= m;
            float vn = fmaf(V[j], c.va, c.vb);
            vn = fmaf(gs, c.vs, vn);
            V[j] = fabsf(vn);
        }
    }
    double acc = 0.0;
#pragma unroll
    for (int j = 0; j < P; ++j) acc += (double)fmaxf(S[j] - c.K, 0.0f) + 1e-30 * x[j].v4;
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}

template <int P, int THREADS, int MINB, int UNROLL> void run(const char *name, const Consts &c, double *d_out)
{
    const unsigned n = 1u << 24, blocks = n / (P * THREADS);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaFuncAttributes attr; cudaFuncGetAttributes(&attr, fe<P, THREADS, MINB, UNROLL>);
    float best = 1e30f; double sum = 0;
    for (int rep = 0; rep < 4; ++rep) {
        cudaMemset(d_out, 0, 8);
        cudaEventRecord(e0); fe<P, THREADS, MINB, UNROLL><<<blocks, THREADS>>>(c, d_out); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        cudaMemcpy(&sum, d_out, 8, cudaMemcpyDeviceToHost);
    }
    printf("%-40s P=%d T=%3d minb=%2d regs=%3d  %7.3f ms  %.4e path-steps/s  E=%.6f\n", name, P, THREADS, MINB, attr.numRegs, best,
           (double)n * c.N / (best * 1e-3), sum / n);
}

int main()
{
    Consts c;
    const float dt = 1e-3f, c0 = 1.17741002f, rho = -0.7f;
    c.crdt = 1.0f; c.zr = rho * sqrtf(dt) * c0; c.zc = sqrtf(1 - rho * rho) * sqrtf(dt) * c0;
    c.va = 1.0f - 0.5f * dt; c.vb = 0.5f * 0.1f * dt; c.vs = 0.3f * sqrtf(dt) * c0;
    c.S0 = 1.0f; c.v0 = 0.1f; c.K = 1.0f; c.N = 1000;
    double *d_out; cudaMalloc(&d_out, 8);
    run<1, 256, 8, 5>("XORWOW stream + native step", c, d_out);
    run<2, 128, 12, 5>("XORWOW stream + native step", c, d_out);
    run<2, 128, 16, 5>("XORWOW stream + native step", c, d_out);
    run<4, 128, 8, 5>("XORWOW stream + native step", c, d_out);
    run<4, 128, 10, 5>("XORWOW stream + native step", c, d_out);
    run<1, 128, 16, 5>("XORWOW stream + native step", c, d_out);
    run<2, 128, 12, 1>("... rolled loop (register moves)", c, d_out);
    return 0;
}
