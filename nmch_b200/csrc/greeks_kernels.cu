// greeks_kernels.cu -- strike vector with pathwise delta AND pathwise vega (d price / d v_0) in one pass
// (SURVEY.md §8f rank 2 "a strike vector and pathwise delta/vega reuse the same paths at ~zero RNG cost"; the
// reference fixes K = S_0 and has no sensitivities, src/NMCH/methods/NMCH.cu:7).
//
// fe_tangent_kernel: the native FE step on the native Philox stream (the words, the instructions and therefore the
//   (S_T, V_T) of fe_philox_kernel, bit for bit) carrying the tangent (dV/dv_0, dS/dv_0) of every path in registers;
//   leaves S_T, V_T and dS_T/dv_0 in HBM (12 bytes per path) and reduces the K = S_0 payoff moments like every path kernel.
// strike_greeks_kernel: folds the two terminal arrays per strike: payoff moments, delta = 1{S_T > K} S_T / S_0,
//   vega = 1{S_T > K} dS_T/dv_0 and its square (for the standard error), through the deterministic ticket reduction.
// Not a throughput path (one path per thread, no hoisting): 2^24 paths x 1000 steps take about twice the plain pass.
#include "fe_step.cuh"
#include "kernels.cuh"

namespace nmchb {

constexpr int kTangentThreads = 256;

template <int FLOOR>
__global__ void __launch_bounds__(kTangentThreads)
fe_tangent_kernel(const __grid_constant__ FeLaunch L, ReduceBuffers rb, float *__restrict__ S_out,
                  float *__restrict__ V_out, float *__restrict__ B_out)
{
    const unsigned long long idx = (unsigned long long)blockIdx.x * kTangentThreads + threadIdx.x;
    const bool active = idx < L.n_local;
    const unsigned long long g = L.first_path + (active ? idx : 0ull);
    const uint32_t path_lo = (uint32_t)g, path_hi = (uint32_t)(g >> 32);
    NMCHB_ASSERT(blockDim.x == kTangentThreads && (int)blockIdx.x < L.blocks_per_point && L.n_points == 1);
    float S = L.S0, V = L.v0, A = 1.0f, B = 0.0f;
    U4 w{0u, 0u, 0u, 0u};
    for (int n = 0; n < L.N; ++n) {
        const unsigned long long pos = L.draw_offset + 2ull * (unsigned long long)n;    // u32 words consumed so far
        const bool second = (pos & 2ull) != 0ull;                                      // second half of its block
        if (n == 0 || !second) {
            const unsigned long long blk = pos >> 2;
            w = philox4x32_10((uint32_t)blk, (uint32_t)(blk >> 32), path_lo, path_hi, L.keys);
        }
        fe_step_native_tangent<FLOOR>(S, V, A, B, second ? w.z : w.x, second ? w.w : w.y, L.rdt, L.zr, L.zc, L.pt0);
    }
    double pay = 0.0, pay2 = 0.0;
    if (active) {
        pay = payoff_or_nan(S, L.K);
        pay2 = pay * pay;
        S_out[idx] = S;
        V_out[idx] = V;
        B_out[idx] = B;
    }
    block_reduce_and_finish(pay, pay2, rb.partials, rb.tickets, rb.out, 0, blockIdx.x, L.blocks_per_point);
}

constexpr int kGreekThreads = 256;
constexpr int kGreekBlocks = 296;           // 2 per SM, contiguous slices, 16-byte loads (as strike_moments_kernel)

__device__ __forceinline__ void greek_fold(float s, float b, float K, float inv_S0, double &pay, double &pay2,
                                           double &delta, double &itm, double &vega, double &vega2)
{
    const float p = fmaxf(s - K, 0.0f);
    pay += (double)p;
    pay2 += (double)p * (double)p;
    if (s > K) {
        delta += (double)(s * inv_S0);
        itm += 1.0;
        vega += (double)b;
        vega2 += (double)b * (double)b;
    }
}

__global__ void __launch_bounds__(kGreekThreads)
strike_greeks_kernel(const float *__restrict__ S, const float *__restrict__ B, unsigned long long n,
                     const float *__restrict__ strikes, float inv_S0, ReduceBuffers rb)
{
    const int j = blockIdx.y;
    const float K = strikes[j];
    double pay = 0.0, pay2 = 0.0, delta = 0.0, itm = 0.0, vega = 0.0, vega2 = 0.0;
    const unsigned long long n4 = n / 4ull;
    const float4 *S4 = reinterpret_cast<const float4 *>(S);
    const float4 *B4 = reinterpret_cast<const float4 *>(B);
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const float4 s = __ldg(S4 + i), b = __ldg(B4 + i);
        greek_fold(s.x, b.x, K, inv_S0, pay, pay2, delta, itm, vega, vega2);
        greek_fold(s.y, b.y, K, inv_S0, pay, pay2, delta, itm, vega, vega2);
        greek_fold(s.z, b.z, K, inv_S0, pay, pay2, delta, itm, vega, vega2);
        greek_fold(s.w, b.w, K, inv_S0, pay, pay2, delta, itm, vega, vega2);
    }
    if (blockIdx.x == 0)                                     // ragged tail
        for (unsigned long long i = n4 * 4ull + threadIdx.x; i < n; i += blockDim.x)
            greek_fold(S[i], B[i], K, inv_S0, pay, pay2, delta, itm, vega, vega2);
    block_reduce_and_finish(pay, pay2, rb.partials, rb.tickets, rb.out, 3 * j, blockIdx.x, gridDim.x);
    block_reduce_and_finish(delta, itm, rb.partials, rb.tickets, rb.out, 3 * j + 1, blockIdx.x, gridDim.x);
    block_reduce_and_finish(vega, vega2, rb.partials, rb.tickets, rb.out, 3 * j + 2, blockIdx.x, gridDim.x);
}

cudaError_t launch_fe_tangent(const FeLaunch &L, int floor_kind, ReduceBuffers rb, float *S_out, float *V_out, float *B_out,
                              cudaStream_t stream, KernelInfo *info)
{
    dim3 grid((unsigned)L.blocks_per_point, 1, 1);
    if (floor_kind == kFloorAbs)
        fe_tangent_kernel<kFloorAbs><<<grid, kTangentThreads, 0, stream>>>(L, rb, S_out, V_out, B_out);
    else
        fe_tangent_kernel<kFloorPlus><<<grid, kTangentThreads, 0, stream>>>(L, rb, S_out, V_out, B_out);
    if (info) {
        info->grid_x = (int)grid.x;
        info->grid_y = 1;
        info->block_threads = kTangentThreads;
        info->paths_per_thread = 1;
        cudaFuncAttributes fa{};
        const void *fn = floor_kind == kFloorAbs ? (const void *)fe_tangent_kernel<kFloorAbs> : (const void *)fe_tangent_kernel<kFloorPlus>;
        if (cudaFuncGetAttributes(&fa, fn) == cudaSuccess) info->regs_per_thread = fa.numRegs;
        info->param_bytes = (int)(sizeof(FeLaunch) + sizeof(ReduceBuffers) + 3 * sizeof(float *));
    }
    return cudaGetLastError();
}

cudaError_t launch_strike_greeks(const float *d_S, const float *d_B, unsigned long long n_local, const float *d_strikes,
                                 int n_strikes, float S0, ReduceBuffers rb, cudaStream_t stream)
{
    dim3 grid(kGreekBlocks, (unsigned)n_strikes, 1);
    strike_greeks_kernel<<<grid, kGreekThreads, 0, stream>>>(d_S, d_B, n_local, d_strikes, 1.0f / S0, rb);
    return cudaGetLastError();
}

int greek_blocks_per_slot() { return kGreekBlocks; }

}  // namespace nmchb
