// fe_variants.cu -- standalone throughput experiments on variants of the native FE step loop (not product code).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I ../../nmch_b200/csrc -o fe_variants fe_variants.cu
// Each variant simulates 2^24 paths x 1000 steps and prints path-steps/s and the mean payoff (sanity).
#include <cstdio>
#include <cuda_runtime.h>

#include "device_common.cuh"

using namespace nmchb;

struct Consts {
    PhiloxKeys keys;
    float crdt, zr, zc, va, vb, vs, S0, v0, K;
    int N;
};

template <int MODE>
__device__ __forceinline__ void step(float &S, float &V, uint32_t wa, uint32_t wb, const Consts &c)
{
    const float f1 = bits_to_1_2(wa);
    const float f2 = bits_to_1_2(wb);
    const float u = f1 - 0.99999994f;
    const float l2 = lg2_approx(u);
    const float q = sqrt_approx(-(V * l2));
    const float ang = f2 * 6.2831855f;
    const float gs = q * sin_approx(ang);
    const float gc = q * cos_approx(ang);
    float m = fmaf(gs, c.zr, c.crdt);
    m = fmaf(gc, c.zc, m);
    S *= m;
    float vn = fmaf(V, c.va, c.vb);
    vn = fmaf(gs, c.vs, vn);
    if (MODE == 0) V = fabsf(vn);
    else if (MODE == 1) V = fmaxf(vn, 0.0f);
    else V = __uint_as_float(__float_as_uint(vn) & 0x7fffffffu);   // abs on the ALU pipe
}

// uniform from already-spliced float bits (dense variant)
template <int MODE>
__device__ __forceinline__ void step_f(float &S, float &V, float f1, float f2, const Consts &c)
{
    const float u = f1 - 0.99999994f;
    const float l2 = lg2_approx(u);
    const float q = sqrt_approx(-(V * l2));
    const float ang = f2 * 6.2831855f;
    const float gs = q * sin_approx(ang);
    const float gc = q * cos_approx(ang);
    float m = fmaf(gs, c.zr, c.crdt);
    m = fmaf(gc, c.zc, m);
    S *= m;
    float vn = fmaf(V, c.va, c.vb);
    vn = fmaf(gs, c.vs, vn);
    V = (MODE == 1) ? fmaxf(vn, 0.0f) : fabsf(vn);
}

// VARIANT 0: product loop (2 steps / block).  1: dense (5 steps / 2 blocks).  2: product loop unrolled x2.
template <int P, int THREADS, int MINB, int VARIANT, int MODE>
__global__ void __launch_bounds__(THREADS, MINB) fe(const __grid_constant__ Consts c, double *out)
{
    const uint32_t path0 = blockIdx.x * (P * THREADS) + threadIdx.x;
    float S[P], V[P];
#pragma unroll
    for (int j = 0; j < P; ++j) { S[j] = c.S0; V[j] = c.v0; }
    uint32_t blk = 0;
    if (VARIANT == 0) {
#pragma unroll 1
        for (int it = 0; it < c.N / 2; ++it) {
#pragma unroll
            for (int j = 0; j < P; ++j) {
                const U4 w = philox4x32_10(blk, 0u, path0 + j * THREADS, 0u, c.keys);
                step<MODE>(S[j], V[j], w.x, w.y, c);
                step<MODE>(S[j], V[j], w.z, w.w, c);
            }
            ++blk;
        }
    } else if (VARIANT == 2) {
#pragma unroll 2
        for (int it = 0; it < c.N / 2; ++it) {
#pragma unroll
            for (int j = 0; j < P; ++j) {
                const U4 w = philox4x32_10(blk, 0u, path0 + j * THREADS, 0u, c.keys);
                step<MODE>(S[j], V[j], w.x, w.y, c);
                step<MODE>(S[j], V[j], w.z, w.w, c);
            }
            ++blk;
        }
    } else if (VARIANT == 3) {
#pragma unroll 1
        for (int it = 0; it < c.N / 3; ++it) {
#pragma unroll
            for (int j = 0; j < P; ++j) {
                const U4 w = philox4x32_10(blk, 0u, path0 + j * THREADS, 0u, c.keys);
                const uint32_t rA = ((w.x >> 9) & 0x7ffffeu) | 0x3f800000u;
                const uint32_t aA = (__funnelshift_r(w.y, w.x, 19) & 0x7ffff8u) | 0x3f800000u;
                const uint32_t rB = ((w.y << 1) & 0x7ffffeu) | 0x3f800000u;
                const uint32_t aB = ((w.z >> 9) & 0x7ffff8u) | 0x3f800000u;
                const uint32_t rC = (__funnelshift_r(w.w, w.z, 21) & 0x7ffffeu) | 0x3f800000u;
                const uint32_t aC = ((w.w << 1) & 0x7ffff8u) | 0x3f800000u;
                step_f<MODE>(S[j], V[j], __uint_as_float(rA), __uint_as_float(aA), c);
                step_f<MODE>(S[j], V[j], __uint_as_float(rB), __uint_as_float(aB), c);
                step_f<MODE>(S[j], V[j], __uint_as_float(rC), __uint_as_float(aC), c);
            }
            ++blk;
        }
    } else if (VARIANT == 4 || VARIANT == 5 || VARIANT == 6 || VARIANT == 7) {
        // 4: hand-hoisted product loop (what fe_philox_kernel runs).  5: the same with the two multiplies that do not
        // depend on the path (M0 * block_lo and M1 * (hi ^ path_hi ^ k1[0])) read from a per-block shared-memory table
        // instead of being recomputed by every thread.  6 / 7: the same pair for the 3-steps-per-block dense loop.
        constexpr bool TABLE = (VARIANT == 5 || VARIANT == 7);
        constexpr bool DENSE = (VARIANT >= 6);
        const int iters = DENSE ? c.N / 3 : c.N / 2;
        __shared__ uint4 tab[TABLE ? 512 : 1];
        if (TABLE) {
            for (int i = threadIdx.x; i < iters; i += THREADS) {
                const unsigned long long sv = (unsigned long long)kPhiloxM0 * (uint32_t)i;
                const uint32_t c2 = (uint32_t)(sv >> 32) ^ 0u ^ c.keys.k1[0];
                const unsigned long long p1 = (unsigned long long)kPhiloxM1 * c2;
                tab[i] = make_uint4((uint32_t)(p1 >> 32) ^ c.keys.k0[1], (uint32_t)p1, (uint32_t)sv ^ c.keys.k1[1], 0u);
            }
            __syncthreads();
        }
        PhiloxPathInv inv[P];
#pragma unroll
        for (int j = 0; j < P; ++j) inv[j] = philox_path_invariants(0u, path0 + j * THREADS, c.keys);
#pragma unroll 1
        for (int it = 0; it < iters; ++it) {
            uint32_t e0, e1, e2;
            if (TABLE) {
                const uint4 t = tab[it];
                e0 = t.x; e1 = t.y; e2 = t.z;
            } else {
                const unsigned long long sv = (unsigned long long)kPhiloxM0 * (uint32_t)it;
                const uint32_t c2 = (uint32_t)(sv >> 32) ^ 0u ^ c.keys.k1[0];
                const unsigned long long p1 = (unsigned long long)kPhiloxM1 * c2;
                e0 = (uint32_t)(p1 >> 32) ^ c.keys.k0[1]; e1 = (uint32_t)p1; e2 = (uint32_t)sv ^ c.keys.k1[1];
            }
#pragma unroll
            for (int j = 0; j < P; ++j) {
                uint32_t c0 = e0 ^ inv[j].lo1, c1 = e1, c2 = inv[j].q_hi ^ e2, c3 = inv[j].q_lo;
#pragma unroll
                for (int r = 2; r < 10; ++r) {
                    const unsigned long long a = (unsigned long long)kPhiloxM0 * c0;
                    const unsigned long long b = (unsigned long long)kPhiloxM1 * c2;
                    const uint32_t n0 = (uint32_t)(b >> 32) ^ c1 ^ c.keys.k0[r];
                    const uint32_t n2 = (uint32_t)(a >> 32) ^ c3 ^ c.keys.k1[r];
                    c1 = (uint32_t)b; c3 = (uint32_t)a; c0 = n0; c2 = n2;
                }
                if (DENSE) {
                    const uint32_t rA = (c0 >> 9) | 0x3f800000u;
                    const uint32_t aA = (__funnelshift_r(c1, c0, 18) & 0x7ffff0u) | 0x3f800000u;
                    const uint32_t rB = (__funnelshift_r(c2, c1, 31) & 0x7fffffu) | 0x3f800000u;
                    const uint32_t aB = ((c2 >> 8) & 0x7ffff0u) | 0x3f800000u;
                    const uint32_t rC = (__funnelshift_r(c3, c2, 21) & 0x7fffffu) | 0x3f800000u;
                    const uint32_t aC = ((c3 << 2) & 0x7ffff0u) | 0x3f800000u;
                    step_f<MODE>(S[j], V[j], __uint_as_float(rA), __uint_as_float(aA), c);
                    step_f<MODE>(S[j], V[j], __uint_as_float(rB), __uint_as_float(aB), c);
                    step_f<MODE>(S[j], V[j], __uint_as_float(rC), __uint_as_float(aC), c);
                } else {
                    step<MODE>(S[j], V[j], c0, c1, c);
                    step<MODE>(S[j], V[j], c2, c3, c);
                }
            }
        }
    } else {
#pragma unroll 1
        for (int it = 0; it < c.N / 5; ++it) {
#pragma unroll
            for (int j = 0; j < P; ++j) {
                const U4 a = philox4x32_10(blk, 0u, path0 + j * THREADS, 0u, c.keys);
                const U4 b = philox4x32_10(blk + 1, 0u, path0 + j * THREADS, 0u, c.keys);
                // 8 uniforms from the top 23 bits, 2 more from the 9 low bits of 3 words each (27 -> 23 bits)
                const uint32_t e0 = ((a.x & 0x1ffu) << 14) | ((a.y & 0x1ffu) << 5) | ((a.z & 0x1ffu) >> 4);
                const uint32_t e1 = ((b.x & 0x1ffu) << 14) | ((b.y & 0x1ffu) << 5) | ((b.z & 0x1ffu) >> 4);
                step<MODE>(S[j], V[j], a.x, a.y, c);
                step<MODE>(S[j], V[j], a.z, a.w, c);
                step<MODE>(S[j], V[j], b.x, b.y, c);
                step<MODE>(S[j], V[j], b.z, b.w, c);
                step_f<MODE>(S[j], V[j], __uint_as_float(e0 | 0x3f800000u), __uint_as_float(e1 | 0x3f800000u), c);
            }
            blk += 2;
        }
    }
    double acc = 0.0;
#pragma unroll
    for (int j = 0; j < P; ++j) acc += (double)fmaxf(S[j] - c.K, 0.0f);
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}

template <int P, int THREADS, int MINB, int VARIANT, int MODE>
void run(const char *name, const Consts &c, double *d_out)
{
    const unsigned n = 1u << 24;
    const unsigned blocks = n / (P * THREADS);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaFuncAttributes attr;
    cudaFuncGetAttributes(&attr, fe<P, THREADS, MINB, VARIANT, MODE>);
    float best = 1e30f;
    double sum = 0;
    for (int rep = 0; rep < 4; ++rep) {
        cudaMemset(d_out, 0, 8);
        cudaEventRecord(e0);
        fe<P, THREADS, MINB, VARIANT, MODE><<<blocks, THREADS>>>(c, d_out);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
        cudaMemcpy(&sum, d_out, 8, cudaMemcpyDeviceToHost);
    }
    printf("%-46s P=%d T=%3d minb=%d regs=%3d  %7.3f ms  %.4e path-steps/s  E=%.6f\n", name, P, THREADS, MINB, attr.numRegs, best,
           (double)n * c.N / (best * 1e-3), sum / n);
}

int main()
{
    Consts c;
    c.keys = philox_expand_keys(1234);
    const float dt = 1e-3f, c0 = 1.17741002f, rho = -0.7f;
    c.crdt = 1.0f; c.zr = rho * sqrtf(dt) * c0; c.zc = sqrtf(1 - rho * rho) * sqrtf(dt) * c0;
    c.va = 1.0f - 0.5f * dt; c.vb = 0.5f * 0.1f * dt; c.vs = 0.3f * sqrtf(dt) * c0;
    c.S0 = 1.0f; c.v0 = 0.1f; c.K = 1.0f; c.N = 1000;
    double *d_out;
    cudaMalloc(&d_out, 8);
    c.N = 999;
    run<4, 128, 10, 0, 0>("product loop (N=998), compiler-hoisted", c, d_out);
    run<4, 128, 10, 4, 0>("product loop, hand-hoisted", c, d_out);
    run<4, 128, 10, 5, 0>("product loop, shared multiplies from smem table", c, d_out);
    run<8, 128, 6, 5, 0>("... table, P=8 minb 6", c, d_out);
    run<2, 128, 12, 5, 0>("... table, P=2 minb 12", c, d_out);
    run<4, 128, 10, 6, 0>("3 steps per block (23/19-bit fields), hand-hoisted", c, d_out);
    run<4, 128, 10, 7, 0>("3 steps per block, smem table", c, d_out);
    run<2, 128, 12, 7, 0>("3 steps per block, smem table, P=2 minb 12", c, d_out);
    run<2, 128, 12, 6, 0>("3 steps per block, hand-hoisted, P=2 minb 12", c, d_out);
    return 0;
}
