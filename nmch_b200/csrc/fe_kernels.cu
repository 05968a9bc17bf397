// fe_kernels.cu -- forward-Euler Heston path kernels for sm_100a.
//
// Replaces FE_k1 / FE_k2 / FE_k2_philox / FE_k3 of the reference (src/NMCH/methods/NMCH_FE.cu:16-304)
// with ONE native kernel (fe_philox_kernel; its EXACT instantiation is the cuRAND-Philox-compatible validation
// mode) plus a validation kernel for the sequential XORWOW stream (fe_compat_kernel); both validation paths
// reproduce the reference's cuRAND draws and IEEE arithmetic draw for draw.
//
// Native kernel, per path-step (see DESIGN.md §Kernels for the instruction budget):
//   * half a Philox4x32-10 block (the block serves two steps; round keys sit in uniform registers, the
//     counter-invariant part of rounds 1-2 is hoisted per path)
//   * u32 -> float with I2FP (ALU pipe; not the XU-pipe I2F) + one FFMA: cuRAND's own uniforms
//   * Box-Muller with MUFU.LG2 / MUFU.SIN / MUFU.COS and ONE MUFU.SQRT shared with the SDE:
//       sqrt(V)*sqrt(-2 ln u) = c0 * sqrt(-V * lg2 u),  c0 = sqrt(2 ln 2) folded into host constants
//   * S' = S + S * (r dt + q sin * zr + q cos * zc),  V' = g(V*va + vb + q sin * vs)
//   state (S, V, counters) stays in registers for all N steps: zero HBM traffic in the loop.
#include "compat_math.cuh"
#include "fe_step.cuh"
#include "kernels.cuh"
#include "xorwow_device.cuh"

namespace nmchb {

// ------------------------------------------------------------------------------------------
// native Philox kernel
// ------------------------------------------------------------------------------------------
// One step in either arithmetic: the fused fast-math step above, or (EXACT) cuRAND's IEEE Box-Muller and the
// reference's pinned update on the same two Philox words -- the Philox-compatible validation mode, which thereby
// shares the counter hoisting, the tiling and the several-paths-per-thread structure of the product kernel.
template <int FLOOR, bool EXACT>
__device__ __forceinline__ void fe_step_any(float &S, float &V, uint32_t wa, uint32_t wb, const FeLaunch &L,
                                            const FePoint &pc)
{
    if constexpr (EXACT) {
        float gx, gy;
        box_muller_compat(wa, wb, gx, gy);
        // for EXACT launches the point record carries the raw (k, theta, sigma) in (va, vb, vs)
        fe_step_compat<FLOOR>(S, V, gx, gy, L.r, pc.va, L.rho, pc.vb, pc.vs, L.dt, L.sqrt_dt, L.sqrt_rho);
    } else {
        fe_step_native<FLOOR>(S, V, wa, wb, L.rdt, L.zr, L.zc, pc);
    }
}

// Resident blocks per SM the register allocator must leave room for (measured sweep, profiles/r01_fe_variants.txt):
// 48 registers at P = 4 keeps 40 warps per SM in flight, which beats both more ILP and more occupancy.
template <int P, int THREADS, bool EXACT = false>
struct FeOccupancy {
#ifndef NMCHB_FE_WARPS_P4
#define NMCHB_FE_WARPS_P4 40
#endif
    // P = 1 (30 registers) is what small launches get (engine.cu pick_paths_per_thread): 56 warps per SM = 14 blocks of
    // 128 paths keep BASELINE configs[0] (2^18 paths = 13.8 blocks per SM) in ONE wave; with 12 resident blocks the
    // last 272 of its 2048 blocks ran as a second, nearly empty wave (round 1: 0.66 of the roofline at that size).
    static constexpr int kWarpsTarget = EXACT ? 32 : (P == 1 ? 56 : (P == 2 ? 48 : (P == 4 ? NMCHB_FE_WARPS_P4 : 24)));
    static constexpr int kMinBlocks = kWarpsTarget * 32 / THREADS;
};

template <int P, int FLOOR, int THREADS, bool EXACT>
__global__ void __launch_bounds__(THREADS, FeOccupancy<P, THREADS, EXACT>::kMinBlocks)
fe_philox_kernel(const __grid_constant__ FeLaunch L, const FePoint *__restrict__ pts, ReduceBuffers rb,
                 float *__restrict__ S_out, float *__restrict__ V_out)
{
    constexpr int TILE = P * THREADS;
    const int point = blockIdx.y;
    NMCHB_ASSERT(blockDim.x == THREADS && point < L.n_points && (int)blockIdx.x < L.blocks_per_point);
    NMCHB_ASSERT(L.first_path % TILE == 0 && L.tiles_per_block >= 1);
    const FePoint pc = (pts != nullptr) ? pts[point] : (EXACT ? FePoint{L.raw0.k, L.raw0.theta, L.raw0.sigma, 0.0f} : L.pt0);

    // stream position of this point: what `point` sequential compute() calls would have consumed
    const unsigned long long w0 = L.draw_offset + (unsigned long long)point * 2ull * (unsigned long long)L.N;
    const bool half_start = (w0 & 2ull) != 0ull;

    // per-thread FP64 payoff sums live in shared memory between tiles, not in registers across the step loop
    __shared__ double2 s_acc[THREADS];
    s_acc[threadIdx.x] = make_double2(0.0, 0.0);
    for (int t = 0; t < L.tiles_per_block; ++t) {
        const unsigned long long tile = (unsigned long long)blockIdx.x * L.tiles_per_block + t;
        const unsigned long long local0 = tile * TILE;
        if (local0 >= L.n_local) break;
        const unsigned long long g0 = L.first_path + local0;     // multiple of TILE: no carry below
        uint32_t path_hi = (uint32_t)(g0 >> 32);
        asm volatile("" : "+r"(path_hi));                         // one register, not re-derived from the tile index in the loop
        const uint32_t path_lo0 = (uint32_t)g0 + threadIdx.x;
        NMCHB_ASSERT(((g0 + (unsigned long long)(TILE - 1)) >> 32) == (g0 >> 32));   // the tile's paths share path_hi

        float S[P], V[P];
#pragma unroll
        for (int j = 0; j < P; ++j) {
            S[j] = L.S0;
            V[j] = L.v0;
        }
        unsigned long long blk = w0 >> 2;
        int n = L.N;
        const bool resume = half_start && n > 0;                  // resume in the middle of a block
        if (resume) {
#pragma unroll
            for (int j = 0; j < P; ++j) {
                const U4 w = philox4x32_10((uint32_t)blk, (uint32_t)(blk >> 32), path_lo0 + j * THREADS, path_hi, L.keys);
                fe_step_any<FLOOR, EXACT>(S[j], V[j], w.z, w.w, L, pc);
            }
        }
        blk += resume ? 1ull : 0ull;                              // launch-uniform counters stay outside the branch
        n -= resume ? 1 : 0;
        // Full blocks, two steps each.  The loop is split where the low counter word would wrap so that the
        // high word is loop-invariant: Philox round 1 and the per-path multiply of round 2 then depend only on
        // (path, high word) and are hoisted; the multiplies on the low word are shared by the thread's P paths.
        int pairs = n >> 1;
        while (pairs > 0) {
            const uint32_t blk_hi = (uint32_t)(blk >> 32);
            const uint32_t blk_lo = (uint32_t)blk;
            const unsigned long long room = 0x100000000ull - (unsigned long long)blk_lo;
            const int chunk = ((unsigned long long)pairs < room) ? pairs : (int)room;
            PhiloxPathInv inv[P];
#pragma unroll
            for (int j = 0; j < P; ++j) inv[j] = philox_path_invariants(blk_hi, path_lo0 + j * THREADS, L.keys);
#pragma unroll 1
            for (int it = 0; it < chunk; ++it) {
                const PhiloxBlockUniform bu = philox_block_uniform(blk_lo + (uint32_t)it, path_hi, L.keys);   // uniform datapath
#pragma unroll
                for (int j = 0; j < P; ++j) {
                    const U4 w = philox4x32_10_hoisted(bu, inv[j], L.keys);
                    fe_step_any<FLOOR, EXACT>(S[j], V[j], w.x, w.y, L, pc);
                    fe_step_any<FLOOR, EXACT>(S[j], V[j], w.z, w.w, L, pc);
                }
            }
            blk += (unsigned long long)chunk;
            pairs -= chunk;
        }
        if (n & 1) {
#pragma unroll
            for (int j = 0; j < P; ++j) {
                const U4 w = philox4x32_10((uint32_t)blk, (uint32_t)(blk >> 32), path_lo0 + j * THREADS, path_hi, L.keys);
                fe_step_any<FLOOR, EXACT>(S[j], V[j], w.x, w.y, L, pc);
            }
        }
        double2 acc = s_acc[threadIdx.x];
#pragma unroll
        for (int j = 0; j < P; ++j) {
            const unsigned long long idx = local0 + (unsigned long long)(j * THREADS) + threadIdx.x;
            if (idx < L.n_local) {
                const double pay = payoff_or_nan(S[j], L.K);
                acc.x += pay;
                acc.y += pay * pay;
                if (S_out != nullptr && point == L.n_points - 1) {
                    S_out[idx] = S[j];
                    V_out[idx] = V[j];
                }
            }
        }
        s_acc[threadIdx.x] = acc;
    }
    const double2 total = s_acc[threadIdx.x];
    block_reduce_and_finish(total.x, total.y, rb.partials, rb.tickets, rb.out, point, blockIdx.x, L.blocks_per_point);
}

// ------------------------------------------------------------------------------------------
// dense-draw variant (opt-in stream mode NMCH_RNG_PHILOX_DENSE): THREE steps per Philox block.
// The product kernel is bound by the 32x32->64 multiplies of Philox; this variant spends a third
// fewer of them by cutting each 128-bit block into three (23-bit radius, 19-bit angle) field pairs instead of two
// (32, 32) word pairs.  The radius uniform keeps the 23 bits of the native mode (same tails); only the angle is
// coarser, and an equispaced angle grid of 2^19 points leaves every trigonometric moment below that order exact.
// The price: a (path, step) no longer consumes the words cuRAND's layout assigns to it, so this mode is checked
// against a restatement of ITS mapping and statistically, not against the reference's Philox stream.
//   step A: radius = x[31:9]               angle = x[8:0]  : y[31:22]
//   step B: radius = y[21:0] : z[31]       angle = z[30:12]
//   step C: radius = z[11:0] : w[31:21]    angle = w[20:2]                 (2 bits unused)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void dense_fields(const U4 &w, int phase, float &f1, float &f2)
{
    uint32_t r, a;                                   // mantissas: 23-bit radius field, angle field << 4
    if (phase == 0) {
        r = w.x >> 9;
        a = __funnelshift_r(w.y, w.x, 18) & 0x7ffff0u;
    } else if (phase == 1) {
        r = __funnelshift_r(w.z, w.y, 31) & 0x7fffffu;
        a = (w.z >> 8) & 0x7ffff0u;
    } else {
        r = __funnelshift_r(w.w, w.z, 21) & 0x7fffffu;
        a = (w.w << 2) & 0x7ffff0u;
    }
    f1 = __uint_as_float(r | 0x3f800000u);
    f2 = __uint_as_float(a | 0x3f800000u);
}

template <int FLOOR>
__device__ __forceinline__ void fe_step_dense(float &S, float &V, const U4 &w, int phase, const FeLaunch &L,
                                              const FePoint &pc)
{
    float f1, f2;
    dense_fields(w, phase, f1, f2);
    const float u = f1 - 0.99999994f;                 // (k23 + 0.5) * 2^-23, in (0,1): exact, as the native mode
    const float l2 = lg2_approx(u);
    const float q = sqrt_approx(-(V * l2));
    const float ang = f2 * 6.2831855f;
    const float gs = q * sin_approx(ang);
    const float gc = q * cos_approx(ang);
    float m = fmaf(gs, L.zr, L.rdt);
    m = fmaf(gc, L.zc, m);
    S = fmaf(S, m, S);
    float vn = fmaf(V, pc.va, pc.vb);
    vn = fmaf(gs, pc.vs, vn);
    V = (FLOOR == kFloorAbs) ? fabsf(vn) : fmaxf(vn, 0.0f);
}

template <int P, int FLOOR, int THREADS>
__global__ void __launch_bounds__(THREADS, FeOccupancy<P, THREADS>::kMinBlocks)
fe_dense_kernel(const __grid_constant__ FeLaunch L, const FePoint *__restrict__ pts, ReduceBuffers rb,
                float *__restrict__ S_out, float *__restrict__ V_out)
{
    constexpr int TILE = P * THREADS;
    const int point = blockIdx.y;
    NMCHB_ASSERT(blockDim.x == THREADS && point < L.n_points && (int)blockIdx.x < L.blocks_per_point);
    NMCHB_ASSERT(L.first_path % TILE == 0 && L.tiles_per_block >= 1 && L.dense_r0 < 3u && L.dense_rN < 3u);
    const FePoint pc = (pts != nullptr) ? pts[point] : L.pt0;
    // stream position in STEPS (draw_offset counts two logical draws per step, like the other modes):
    //   s0 = draw_offset / 2 + point * N = 3 * blk0 + phase0, from the host's pre-divided parts -- a 64-bit division
    //   here would take the block index, and with it the path-independent Philox multiplies, off the uniform datapath
    const unsigned int t0 = L.dense_r0 + (unsigned int)point * L.dense_rN;
    // t0 / 3, exact below 2^17 (t0 <= 2 + 2 * 65534).  The product needs 34 bits: a 32-bit multiply wraps from t0 = 98304
    // on (a sweep of more than 49151 points with N % 3 == 2), so it is formed in 64 bits (one wide multiply + shift, uniform datapath)
    const unsigned int q0 = (unsigned int)(((unsigned long long)t0 * 43691ull) >> 17);
    const unsigned long long blk0 = L.dense_q0 + (unsigned long long)point * (unsigned long long)L.dense_qN + (unsigned long long)q0;
    const int phase0 = (int)(t0 - 3u * q0);
    NMCHB_ASSERT(phase0 >= 0 && phase0 < 3 && q0 == t0 / 3u);

    __shared__ double2 s_acc[THREADS];
    s_acc[threadIdx.x] = make_double2(0.0, 0.0);
    for (int t = 0; t < L.tiles_per_block; ++t) {
        const unsigned long long tile = (unsigned long long)blockIdx.x * L.tiles_per_block + t;
        const unsigned long long local0 = tile * TILE;
        if (local0 >= L.n_local) break;
        const unsigned long long g0 = L.first_path + local0;
        uint32_t path_hi = (uint32_t)(g0 >> 32);
#ifndef NMCHB_DENSE_NO_PIN
        asm volatile("" : "+r"(path_hi));                         // one register, instead of re-deriving it from the tile index every iteration
#endif
        const uint32_t path_lo0 = (uint32_t)g0 + threadIdx.x;
        float S[P], V[P];
#pragma unroll
        for (int j = 0; j < P; ++j) {
            S[j] = L.S0;
            V[j] = L.v0;
        }
        unsigned long long blk = blk0;
        int n = L.N;
        if (phase0 != 0 && n > 0) {                               // finish the block a previous call left half used
            const int cnt = min(3 - phase0, n);
#pragma unroll
            for (int j = 0; j < P; ++j) {
                const U4 w = philox4x32_10((uint32_t)blk, (uint32_t)(blk >> 32), path_lo0 + j * THREADS, path_hi, L.keys);
#pragma unroll
                for (int ph = 1; ph < 3; ++ph)                    // static phases: the field extraction stays branch-free
                    if (ph >= phase0 && ph < phase0 + cnt) fe_step_dense<FLOOR>(S[j], V[j], w, ph, L, pc);
            }
            ++blk;
            n -= cnt;
        }
        int triples = n / 3;
        while (triples > 0) {                                     // same counter split as the product kernel
            const uint32_t blk_hi = (uint32_t)(blk >> 32);
            const uint32_t blk_lo = (uint32_t)blk;
            const unsigned long long room = 0x100000000ull - (unsigned long long)blk_lo;
            const int chunk = ((unsigned long long)triples < room) ? triples : (int)room;
            PhiloxPathInv inv[P];
#pragma unroll
            for (int j = 0; j < P; ++j) inv[j] = philox_path_invariants(blk_hi, path_lo0 + j * THREADS, L.keys);
#pragma unroll 1
            for (int it = 0; it < chunk; ++it) {
                const PhiloxBlockUniform bu = philox_block_uniform(blk_lo + (uint32_t)it, path_hi, L.keys);
#pragma unroll
                for (int j = 0; j < P; ++j) {
                    const U4 w = philox4x32_10_hoisted(bu, inv[j], L.keys);
                    fe_step_dense<FLOOR>(S[j], V[j], w, 0, L, pc);
                    fe_step_dense<FLOOR>(S[j], V[j], w, 1, L, pc);
                    fe_step_dense<FLOOR>(S[j], V[j], w, 2, L, pc);
                }
            }
            blk += (unsigned long long)chunk;
            triples -= chunk;
        }
        const int rest = n % 3;
        if (rest > 0) {
#pragma unroll
            for (int j = 0; j < P; ++j) {
                const U4 w = philox4x32_10((uint32_t)blk, (uint32_t)(blk >> 32), path_lo0 + j * THREADS, path_hi, L.keys);
#pragma unroll
                for (int ph = 0; ph < 2; ++ph)
                    if (ph < rest) fe_step_dense<FLOOR>(S[j], V[j], w, ph, L, pc);
            }
        }
        double2 acc = s_acc[threadIdx.x];
#pragma unroll
        for (int j = 0; j < P; ++j) {
            const unsigned long long idx = local0 + (unsigned long long)(j * THREADS) + threadIdx.x;
            if (idx < L.n_local) {
                const double pay = payoff_or_nan(S[j], L.K);
                acc.x += pay;
                acc.y += pay * pay;
                if (S_out != nullptr && point == L.n_points - 1) {
                    S_out[idx] = S[j];
                    V_out[idx] = V[j];
                }
            }
        }
        s_acc[threadIdx.x] = acc;
    }
    const double2 total = s_acc[threadIdx.x];
    block_reduce_and_finish(total.x, total.y, rb.partials, rb.tickets, rb.out, point, blockIdx.x, L.blocks_per_point);
}

template <int P, int FLOOR>
static cudaError_t launch_dense_t(const FeLaunch &L, const FePoint *d_pts, ReduceBuffers rb, float *S_out, float *V_out,
                                  cudaStream_t stream, KernelInfo *info)
{
    dim3 grid((unsigned)L.blocks_per_point, (unsigned)L.n_points, 1);
    fe_dense_kernel<P, FLOOR, 128><<<grid, 128, 0, stream>>>(L, d_pts, rb, S_out, V_out);
    cudaFuncAttributes attr{};
    const cudaError_t err = cudaFuncGetAttributes(&attr, fe_dense_kernel<P, FLOOR, 128>);
    if (info)
        *info = KernelInfo{(int)grid.x, (int)grid.y, 128, P, attr.numRegs,
                           (int)(sizeof(FeLaunch) + sizeof(const FePoint *) + sizeof(ReduceBuffers) + 2 * sizeof(float *))};
    const cudaError_t lerr = cudaGetLastError();
    return lerr != cudaSuccess ? lerr : err;
}

cudaError_t launch_fe_dense(const FeLaunch &L, int floor_kind, int P, const FePoint *d_pts, ReduceBuffers rb,
                            float *S_out, float *V_out, cudaStream_t stream, KernelInfo *info)
{
    const bool abs_floor = floor_kind == kFloorAbs;
    switch (P) {
    case 1: return abs_floor ? launch_dense_t<1, kFloorAbs>(L, d_pts, rb, S_out, V_out, stream, info)
                             : launch_dense_t<1, kFloorPlus>(L, d_pts, rb, S_out, V_out, stream, info);
    case 2: return abs_floor ? launch_dense_t<2, kFloorAbs>(L, d_pts, rb, S_out, V_out, stream, info)
                             : launch_dense_t<2, kFloorPlus>(L, d_pts, rb, S_out, V_out, stream, info);
    case 4: return abs_floor ? launch_dense_t<4, kFloorAbs>(L, d_pts, rb, S_out, V_out, stream, info)
                             : launch_dense_t<4, kFloorPlus>(L, d_pts, rb, S_out, V_out, stream, info);
    default: return cudaErrorInvalidValue;
    }
}

template <int P, int FLOOR, bool EXACT>
static cudaError_t launch_philox_t(const FeLaunch &L, int block_threads, const FePoint *d_pts,
                                   ReduceBuffers rb, float *S_out, float *V_out, cudaStream_t stream,
                                   KernelInfo *info)
{
    dim3 grid((unsigned)L.blocks_per_point, (unsigned)L.n_points, 1);
    cudaFuncAttributes attr{};
    cudaError_t err;
    if (block_threads == 128) {
        fe_philox_kernel<P, FLOOR, 128, EXACT><<<grid, 128, 0, stream>>>(L, d_pts, rb, S_out, V_out);
        err = cudaFuncGetAttributes(&attr, fe_philox_kernel<P, FLOOR, 128, EXACT>);
    } else {
        fe_philox_kernel<P, FLOOR, 256, EXACT><<<grid, 256, 0, stream>>>(L, d_pts, rb, S_out, V_out);
        err = cudaFuncGetAttributes(&attr, fe_philox_kernel<P, FLOOR, 256, EXACT>);
    }
    if (info) {
        info->grid_x = (int)grid.x;
        info->grid_y = (int)grid.y;
        info->block_threads = block_threads == 128 ? 128 : 256;
        info->paths_per_thread = P;
        info->regs_per_thread = attr.numRegs;
        info->param_bytes = (int)(sizeof(FeLaunch) + sizeof(const FePoint *) + sizeof(ReduceBuffers) + 2 * sizeof(float *));
    }
    const cudaError_t lerr = cudaGetLastError();
    return lerr != cudaSuccess ? lerr : err;
}

cudaError_t launch_fe_philox(const FeLaunch &L, int floor_kind, int P, int block_threads, bool exact_math,
                             const FePoint *d_pts, ReduceBuffers rb, float *S_out, float *V_out, cudaStream_t stream,
                             KernelInfo *info)
{
#define NMCHB_DISPATCH(PP, EX)                                                                                    \
    case PP:                                                                                                      \
        return floor_kind == kFloorAbs                                                                            \
                   ? launch_philox_t<PP, kFloorAbs, EX>(L, block_threads, d_pts, rb, S_out, V_out, stream, info)  \
                   : launch_philox_t<PP, kFloorPlus, EX>(L, block_threads, d_pts, rb, S_out, V_out, stream, info);
    if (exact_math) {
        switch (P) {
            NMCHB_DISPATCH(1, true)
            NMCHB_DISPATCH(2, true)
        default:
            return cudaErrorInvalidValue;
        }
    }
    switch (P) {
        NMCHB_DISPATCH(1, false)
        NMCHB_DISPATCH(2, false)
        NMCHB_DISPATCH(4, false)
        NMCHB_DISPATCH(8, false)
    default:
        return cudaErrorInvalidValue;
    }
#undef NMCHB_DISPATCH
}

// ------------------------------------------------------------------------------------------
// compat kernel: the reference's streams and arithmetic, one path per thread.
// ------------------------------------------------------------------------------------------
struct CompatXorwow {
    uint32_t d, v0, v1, v2, v3, v4;
    __device__ __forceinline__ uint32_t next()
    {   // XORWOW step (published algorithm; cuRAND curand_kernel.h:863-874)
        const uint32_t t = v0 ^ (v0 >> 2);
        v0 = v1; v1 = v2; v2 = v3; v3 = v4;
        v4 = (v4 ^ (v4 << 4)) ^ (t ^ (t << 1));
        d += 362437u;
        return v4 + d;
    }
    // Start of chunk c of a sweep: c * chunk_points * 2N draws further down this path's stream.  FE consumes exactly 2N
    // draws per point, so the position of every point is known in advance: the five xorshift words advance by the
    // GF(2) matrix power A^c (A = T^(2N chunk_points), digit tables from the host), the Weyl word by a multiple.
    __device__ __forceinline__ void skip_to_chunk(const uint32_t *__restrict__ tables, unsigned int c, uint32_t d_per_chunk)
    {
        uint32_t v[5] = {v0, v1, v2, v3, v4};
        xorwow_apply_digits(v, tables, c);
        v0 = v[0]; v1 = v[1]; v2 = v[2]; v3 = v[3]; v4 = v[4];
        d += d_per_chunk * c;
    }
};

template <int FLOOR>
__global__ void __launch_bounds__(256)
fe_compat_kernel(const __grid_constant__ FeLaunch L, const RawPoint *__restrict__ pts, XorwowState xs,
                 const uint32_t *__restrict__ skip, ReduceBuffers rb, float *__restrict__ S_out, float *__restrict__ V_out)
{
    const unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = idx < L.n_local;
    NMCHB_ASSERT((int)gridDim.x == L.blocks_per_point && (unsigned long long)gridDim.x * blockDim.x >= L.n_local);
    NMCHB_ASSERT((int)gridDim.y == L.n_chunks && (L.n_chunks == 1 || skip != nullptr));
    CompatXorwow xw{};
    if (valid) {
        xw.d = xs.d[idx]; xw.v0 = xs.v0[idx]; xw.v1 = xs.v1[idx];
        xw.v2 = xs.v2[idx]; xw.v3 = xs.v3[idx]; xw.v4 = xs.v4[idx];
    }
    // The points of a sweep are walked inside the thread: the XORWOW stream is sequential, and this is exactly the
    // reference's order (one compute() after the other on the state written back by the previous one).  When the
    // paths alone cannot fill the GPU the walk is cut into chunks of consecutive points, blockIdx.y = chunk, each
    // starting from the skipped-ahead state: same draws per (path, point), same blocks and reduction order per point.
    const int chunk = blockIdx.y;
    const int p0 = chunk * L.chunk_points;
    const int p1 = min(p0 + L.chunk_points, L.n_points);
    if (chunk > 0 && valid) xw.skip_to_chunk(skip, (unsigned int)chunk, 362437u * 2u * (uint32_t)L.N * (uint32_t)L.chunk_points);
    for (int point = p0; point < p1; ++point) {
        const RawPoint rp = (pts != nullptr) ? pts[point] : L.raw0;
        float S = L.S0, V = L.v0;
        if (valid) {
            for (int n = 0; n < L.N; ++n) {
                const uint32_t a = xw.next();
                const uint32_t b = xw.next();
                float gx, gy;
                box_muller_compat(a, b, gx, gy);
                fe_step_compat<FLOOR>(S, V, gx, gy, L.r, rp.k, L.rho, rp.theta, rp.sigma, L.dt, L.sqrt_dt,
                                      L.sqrt_rho);
            }
        }
        double pay = 0.0;
        if (valid) {
            pay = (double)fmaxf(0.0f, S - L.K);
            if (S_out != nullptr && point == L.n_points - 1) {
                S_out[idx] = S;
                V_out[idx] = V;
            }
        }
        block_reduce_and_finish(pay, pay * pay, rb.partials, rb.tickets, rb.out, point, blockIdx.x,
                                L.blocks_per_point);
    }
    if (valid && chunk == L.n_chunks - 1) {      // streams continue across compute() calls (NMCH_FE.cu:303): the last
        xs.d[idx] = xw.d; xs.v0[idx] = xw.v0; xs.v1[idx] = xw.v1;     // chunk ends where the whole sweep ends
        xs.v2[idx] = xw.v2; xs.v3[idx] = xw.v3; xs.v4[idx] = xw.v4;
    }
}

// ------------------------------------------------------------------------------------------
// XORWOW stream + native step (opt-in mode NMCH_RNG_XORWOW_FAST): the reference's default generator, and therefore
// the same integer draws per path as its CUDA build on the same seed, but the native arithmetic -- cuRAND's uniforms
// by I2FP + FFMA, MUFU Box-Muller sharing its square root with the SDE, folded constants.  XORWOW needs no
// multiplies (8 ALU-pipe operations per draw), so the FMA pipe that binds the Philox kernels is left to the 11 FP32
// operations of the step: 37.8 cycles per warp-step.  One path per thread (six state words in registers), the step
// loop unrolled by five so that the rotation of the five xorshift words costs no moves; the points of a sweep are
// walked inside the thread because the stream is sequential (the reference's order), state written back at the end.
// ------------------------------------------------------------------------------------------
template <int FLOOR>
__global__ void __launch_bounds__(256, 8)
fe_xorwow_fast_kernel(const __grid_constant__ FeLaunch L, const FePoint *__restrict__ pts, XorwowState xs,
                      const uint32_t *__restrict__ skip, ReduceBuffers rb, float *__restrict__ S_out, float *__restrict__ V_out)
{
    const unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = idx < L.n_local;
    NMCHB_ASSERT((int)gridDim.x == L.blocks_per_point && (unsigned long long)gridDim.x * blockDim.x >= L.n_local);
    NMCHB_ASSERT((int)gridDim.y == L.n_chunks && (L.n_chunks == 1 || skip != nullptr));
    CompatXorwow xw{};
    if (valid) {
        xw.d = xs.d[idx]; xw.v0 = xs.v0[idx]; xw.v1 = xs.v1[idx];
        xw.v2 = xs.v2[idx]; xw.v3 = xs.v3[idx]; xw.v4 = xs.v4[idx];
    }
    const int chunk = blockIdx.y;                        // chunks of consecutive points, as in fe_compat_kernel
    const int p0 = chunk * L.chunk_points;
    const int p1 = min(p0 + L.chunk_points, L.n_points);
    if (chunk > 0 && valid) xw.skip_to_chunk(skip, (unsigned int)chunk, 362437u * 2u * (uint32_t)L.N * (uint32_t)L.chunk_points);
    for (int point = p0; point < p1; ++point) {
        const FePoint pc = (pts != nullptr) ? pts[point] : L.pt0;
        float S = L.S0, V = L.v0;
        if (valid) {
#pragma unroll 5
            for (int n = 0; n < L.N; ++n) {
                const uint32_t a = xw.next();                  // first draw -> radius, second -> angle, x pairs with sin:
                const uint32_t b = xw.next();                  // curand_normal2's order (curand_normal.h:70-87)
                fe_step_native<FLOOR, true>(S, V, a, b, L.rdt, L.zr, L.zc, pc);
            }
        }
        double pay = 0.0;
        if (valid) {
            pay = payoff_or_nan(S, L.K);
            if (S_out != nullptr && point == L.n_points - 1) {
                S_out[idx] = S;
                V_out[idx] = V;
            }
        }
        block_reduce_and_finish(pay, pay * pay, rb.partials, rb.tickets, rb.out, point, blockIdx.x,
                                L.blocks_per_point);
    }
    if (valid && chunk == L.n_chunks - 1) {      // streams continue across compute() calls (NMCH_FE.cu:303)
        xs.d[idx] = xw.d; xs.v0[idx] = xw.v0; xs.v1[idx] = xw.v1;
        xs.v2[idx] = xw.v2; xs.v3[idx] = xw.v3; xs.v4[idx] = xw.v4;
    }
}

cudaError_t launch_fe_xorwow_fast(const FeLaunch &L, int floor_kind, const FePoint *d_pts, XorwowState xs, const uint32_t *skip,
                                  ReduceBuffers rb, float *S_out, float *V_out, cudaStream_t stream, KernelInfo *info)
{
    const int threads = 256;                 // one path per thread; L.blocks_per_point is sized for this
    dim3 grid((unsigned)L.blocks_per_point, (unsigned)L.n_chunks, 1);
    cudaFuncAttributes attr{};
    cudaError_t err = cudaSuccess;
    if (floor_kind == kFloorAbs) {
        fe_xorwow_fast_kernel<kFloorAbs><<<grid, threads, 0, stream>>>(L, d_pts, xs, skip, rb, S_out, V_out);
        err = cudaFuncGetAttributes(&attr, fe_xorwow_fast_kernel<kFloorAbs>);
    } else {
        fe_xorwow_fast_kernel<kFloorPlus><<<grid, threads, 0, stream>>>(L, d_pts, xs, skip, rb, S_out, V_out);
        err = cudaFuncGetAttributes(&attr, fe_xorwow_fast_kernel<kFloorPlus>);
    }
    if (info)
        *info = KernelInfo{(int)grid.x, (int)grid.y, threads, 1, attr.numRegs,
                           (int)(sizeof(FeLaunch) + sizeof(const FePoint *) + sizeof(XorwowState) + sizeof(ReduceBuffers) + 3 * sizeof(float *))};
    const cudaError_t lerr = cudaGetLastError();
    return lerr != cudaSuccess ? lerr : err;
}

cudaError_t launch_fe_compat(const FeLaunch &L, int floor_kind, const RawPoint *d_pts, XorwowState xs, const uint32_t *skip,
                             ReduceBuffers rb, float *S_out, float *V_out, cudaStream_t stream, KernelInfo *info)
{
    const int threads = 256;                 // one path per thread; L.blocks_per_point is sized for this
    dim3 grid((unsigned)L.blocks_per_point, (unsigned)L.n_chunks, 1);
    cudaFuncAttributes attr{};
    cudaError_t err = cudaSuccess;
    if (floor_kind == kFloorAbs) {
        fe_compat_kernel<kFloorAbs><<<grid, threads, 0, stream>>>(L, d_pts, xs, skip, rb, S_out, V_out);
        err = cudaFuncGetAttributes(&attr, fe_compat_kernel<kFloorAbs>);
    } else {
        fe_compat_kernel<kFloorPlus><<<grid, threads, 0, stream>>>(L, d_pts, xs, skip, rb, S_out, V_out);
        err = cudaFuncGetAttributes(&attr, fe_compat_kernel<kFloorPlus>);
    }
    if (info) {
        info->grid_x = (int)grid.x;
        info->grid_y = (int)grid.y;
        info->block_threads = threads;
        info->paths_per_thread = 1;
        info->regs_per_thread = attr.numRegs;
        info->param_bytes = (int)(sizeof(FeLaunch) + sizeof(const RawPoint *) + sizeof(XorwowState) + sizeof(ReduceBuffers) + 2 * sizeof(float *));
    }
    const cudaError_t lerr = cudaGetLastError();
    return lerr != cudaSuccess ? lerr : err;
}

}  // namespace nmchb
