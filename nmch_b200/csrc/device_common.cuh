// device_common.cuh -- Philox4x32-10, approximate-math wrappers and the FP64 payoff reduction
// shared by the FE and EM kernels (sm_100a).
#pragma once
#include <cassert>
#include <cstdint>
#include <cuda_runtime.h>

// Checked build (-DNMCHB_CHECKS, nmch_b200/_build.py build_checked()): device-side asserts on every index the kernels
// form (path, tile, point, partial and ticket slots, state arrays).  A failed assert traps the kernel and the next
// CUDA call of the engine returns NMCH_ERR_CUDA.  Stands in for compute-sanitizer, which is closed on the B200 pool;
// tests/test_gpu_checked_build.py runs the ragged / 64-bit-index / shard / multi-tile cases through that library.
#ifdef NMCHB_CHECKS
#define NMCHB_ASSERT(cond) assert(cond)
#else
#define NMCHB_ASSERT(cond) ((void)0)
#endif

namespace nmchb {

// ---------------------------------------------------------------------------------------
// Philox4x32-10.  Counter layout follows cuRAND (curand_kernel.h:1022-1037): ctr = (block_lo,
// block_hi, path_lo, path_hi), key = seed.  The ten round keys depend only on the seed, so the
// host expands them once and they travel in the kernel parameter block: after unrolling, ptxas
// keeps every key in a uniform register that the consuming LOP3 reads directly -- no per-thread
// key registers, no key additions in the loop.
// ---------------------------------------------------------------------------------------
constexpr uint32_t kPhiloxM0 = 0xD2511F53u;
constexpr uint32_t kPhiloxM1 = 0xCD9E8D57u;
constexpr uint32_t kPhiloxW0 = 0x9E3779B9u;
constexpr uint32_t kPhiloxW1 = 0xBB67AE85u;

struct PhiloxKeys {
    uint32_t k0[10];
    uint32_t k1[10];
};

inline PhiloxKeys philox_expand_keys(unsigned long long seed)
{
    PhiloxKeys k;
    uint32_t a = (uint32_t)seed, b = (uint32_t)(seed >> 32);
    for (int i = 0; i < 10; ++i) {
        k.k0[i] = a;
        k.k1[i] = b;
        a += kPhiloxW0;
        b += kPhiloxW1;
    }
    return k;
}

struct U4 {
    uint32_t x, y, z, w;
};

__device__ __forceinline__ U4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                            const PhiloxKeys &K)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned long long p0 = (unsigned long long)kPhiloxM0 * c0;   // IMAD.WIDE.U32
        const unsigned long long p1 = (unsigned long long)kPhiloxM1 * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ K.k0[r];             // one LOP3
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ K.k1[r];
        c1 = (uint32_t)p1;
        c3 = (uint32_t)p0;
        c0 = n0;
        c2 = n2;
    }
    return U4{c0, c1, c2, c3};
}

// The same block function with the loop-invariant part factored out by hand, for a loop that walks the low
// counter word c0 while (c1, c2, c3) = (block_hi, path_lo, path_hi) stay fixed:
//   round 1:  p1 = M1*c2 is invariant;  p0 = M0*c0 is shared by all paths of the thread
//   round 2:  its c0 = hi(p1)^c1^k0[0] is invariant, so the product M0*c0 is too
// PhiloxPathInv holds the three invariant words of one path; the opaque asm pins them in registers (ptxas
// otherwise rematerialises them every iteration under the register cap: +14 instructions per 8 path-steps).
struct PhiloxPathInv {
    uint32_t lo1;        // lo(M1 * path_lo): c1 entering round 2
    uint32_t q_hi, q_lo; // M0 * (hi(M1*path_lo) ^ block_hi ^ k0[0]): the round-2 multiply on c0
};

__device__ __forceinline__ PhiloxPathInv philox_path_invariants(uint32_t block_hi, uint32_t path_lo, const PhiloxKeys &K)
{
    const unsigned long long p1 = (unsigned long long)kPhiloxM1 * path_lo;
    const uint32_t c0 = (uint32_t)(p1 >> 32) ^ block_hi ^ K.k0[0];
    const unsigned long long q = (unsigned long long)kPhiloxM0 * c0;
    PhiloxPathInv v{(uint32_t)p1, (uint32_t)(q >> 32), (uint32_t)q};
    asm volatile("" : "+r"(v.lo1), "+r"(v.q_hi), "+r"(v.q_lo));
    return v;
}

// The part of a block that depends on neither the path's low word nor the thread: from s = M0 * block_lo and the
// (block-uniform) high path word.  Everything here is the same for every thread of the block, so it belongs on the
// uniform datapath; it is grouped so that each per-path LOP3 of rounds 2-3 sees ONE uniform operand (two uniform
// operands of one LOP3 would force ptxas to keep the whole chain, multiplies included, in vector registers).
struct PhiloxBlockUniform {
    uint32_t c0u;        // hi(M1 * c2') ^ k0[1]      with c2' = hi(s) ^ path_hi ^ k1[0]   -> c0 entering round 2 is c0u ^ inv.lo1
    uint32_t c1u;        // lo(M1 * c2') ^ k0[2]      c1 entering round 2, with round 3's key already folded in
    uint32_t c2u;        // lo(s) ^ k1[1]             c2 entering round 2 is inv.q_hi ^ c2u
};

__device__ __forceinline__ PhiloxBlockUniform philox_block_uniform(uint32_t block_lo, uint32_t path_hi, const PhiloxKeys &K)
{
    const unsigned long long s = (unsigned long long)kPhiloxM0 * block_lo;
    const uint32_t c2_r1 = (uint32_t)(s >> 32) ^ path_hi ^ K.k1[0];
    const unsigned long long p1 = (unsigned long long)kPhiloxM1 * c2_r1;
    return PhiloxBlockUniform{(uint32_t)(p1 >> 32) ^ K.k0[1], (uint32_t)p1 ^ K.k0[2], (uint32_t)s ^ K.k1[1]};
}

__device__ __forceinline__ U4 philox4x32_10_hoisted(const PhiloxBlockUniform &bu, const PhiloxPathInv &inv, const PhiloxKeys &K)
{
    // state entering round 2 (index 1): (c0, c1, c2, c3) = (hi1^lo1'^k0[1], lo(p1), q_hi^lo(s)^k1[1], q_lo)
    uint32_t c0 = bu.c0u ^ inv.lo1;
    uint32_t c2 = inv.q_hi ^ bu.c2u;
    uint32_t c3 = inv.q_lo;
    // round 3 (index 2), with c1 ^ k0[2] pre-folded into the uniform word
    {
        const unsigned long long a = (unsigned long long)kPhiloxM0 * c0;
        const unsigned long long b = (unsigned long long)kPhiloxM1 * c2;
        const uint32_t n0 = (uint32_t)(b >> 32) ^ bu.c1u;
        const uint32_t n2 = (uint32_t)(a >> 32) ^ c3 ^ K.k1[2];
        c3 = (uint32_t)a;
        c0 = n0;
        c2 = n2;
        uint32_t c1 = (uint32_t)b;
#pragma unroll
        for (int r = 3; r < 10; ++r) {
            const unsigned long long a2 = (unsigned long long)kPhiloxM0 * c0;
            const unsigned long long b2 = (unsigned long long)kPhiloxM1 * c2;
            const uint32_t m0 = (uint32_t)(b2 >> 32) ^ c1 ^ K.k0[r];
            const uint32_t m2 = (uint32_t)(a2 >> 32) ^ c3 ^ K.k1[r];
            c1 = (uint32_t)b2;
            c3 = (uint32_t)a2;
            c0 = m0;
            c2 = m2;
        }
        return U4{c0, c1, c2, c3};
    }
}

// s_hi:s_lo = M0 * block_lo (computed once per thread and iteration), path_hi as in the counter.
__device__ __forceinline__ U4 philox4x32_10_hoisted(uint32_t s_hi, uint32_t s_lo, uint32_t path_hi,
                                                    const PhiloxPathInv &inv, const PhiloxKeys &K)
{
    const uint32_t c2_r1 = s_hi ^ path_hi ^ K.k1[0];
    const unsigned long long p1 = (unsigned long long)kPhiloxM1 * c2_r1;
    const PhiloxBlockUniform bu{(uint32_t)(p1 >> 32) ^ K.k0[1], (uint32_t)p1 ^ K.k0[2], s_lo ^ K.k1[1]};
    return philox4x32_10_hoisted(bu, inv, K);
}

// ---------------------------------------------------------------------------------------
// Single-instruction transcendental wrappers (MUFU.*).  .ftz keeps ptxas from wrapping them in
// denormal-scaling sequences; every argument in the kernels is a normal number by construction.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float lg2_approx(float x)
{
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float ex2_approx(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sqrt_approx(float x)
{
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rsqrt_approx(float x)
{
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x)
{
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sin_approx(float x)
{
    float y;
    asm("sin.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float cos_approx(float x)
{
    float y;
    asm("cos.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// u32 -> float in [1,2) from the top 23 bits: SHF + LOP3 on the ALU pipe, no I2F (which shares the
// 16-lane XU pipe with MUFU).  The value equals 1 + floor(x / 2^9) * 2^-23.
__device__ __forceinline__ float bits_to_1_2(uint32_t x)
{
    return __uint_as_float((x >> 9) | 0x3f800000u);
}

// Payoff (S_T - K)^+ of one path as the FP64 summand.  fmaxf returns its non-NaN operand, so a path that went NaN
// (a sampler's hang guard, parameters at the edge of the representable range) would silently count as zero payoff:
// carry the NaN into the sums instead, where the caller sees it.
__device__ __forceinline__ double payoff_or_nan(float S, float K)
{
    const float pay = fmaxf(0.0f, S - K);
    return (S == S) ? (double)pay : (double)S;
}

// (x & mask) | bits as ONE LOP3: written as x & mask | bits with two literals, ptxas emits two LOP3 (one immediate per
// instruction); with the mask in a register (a uniform register after hoisting) it is a single one.  ALU-pipe
// instructions cost two issue cycles each on this part (profiles/r02_pipe_rates2.txt), so this is worth an asm.
__device__ __forceinline__ uint32_t and_or(uint32_t x, uint32_t mask, uint32_t bits)
{
    uint32_t r;
    asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(r) : "r"(x), "r"(mask), "r"(bits));
    return r;
}

// ---------------------------------------------------------------------------------------
// Payoff moments: per-thread FP64 pair -> warp shuffle -> shared memory -> one FP64 partial per
// block -> the last block of the point (atomic ticket) folds the partials in index order with
// Kahan compensation and ONE thread writes the result.  Deterministic for a fixed launch shape.
// Replaces blockReduceSum + 2 float atomicAdd per block (NMCH_FE.cu:85-126, 176-181).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

struct KahanPair {
    double s, c;
    __device__ __forceinline__ void add(double x)
    {
        const double y = x - c;
        const double t = s + y;
        c = (t - s) - y;
        s = t;
    }
};

// Called by every thread of the block.  partials: [n_points][blocks_per_point] double2.
// Returns after the point's final result (if this was the last block) has been written to out[2*point].
__device__ __forceinline__ void block_reduce_and_finish(double a, double b, double2 *partials,
                                                        unsigned int *tickets, double *out, int point,
                                                        int block_in_point, int blocks_per_point)
{
    __shared__ double sh_a[32];
    __shared__ double sh_b[32];
    __shared__ bool sh_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
    a = warp_sum(a);
    b = warp_sum(b);
    __syncthreads();                       // protects sh_* against a previous call in the same kernel
    if (lane == 0) {
        sh_a[warp] = a;
        sh_b[warp] = b;
    }
    __syncthreads();
    if (warp == 0) {
        a = (lane < nwarps) ? sh_a[lane] : 0.0;
        b = (lane < nwarps) ? sh_b[lane] : 0.0;
        a = warp_sum(a);
        b = warp_sum(b);
        if (lane == 0) {
            NMCHB_ASSERT(point >= 0 && block_in_point >= 0 && block_in_point < blocks_per_point);
            partials[(size_t)point * blocks_per_point + block_in_point] = make_double2(a, b);
            __threadfence();
            const unsigned int t = atomicAdd(&tickets[point], 1u);
            sh_last = (t == (unsigned int)blocks_per_point - 1u);
        }
    }
    __syncthreads();
    if (!sh_last) return;
    __threadfence();
    // last block of this point: strided Kahan sums, then a fixed-order fold; thread 0 is the single writer
    const double2 *src = partials + (size_t)point * blocks_per_point;
    KahanPair ka{0.0, 0.0}, kb{0.0, 0.0};
    for (int i = threadIdx.x; i < blocks_per_point; i += blockDim.x) {
        const double2 v = __ldcg(src + i);
        ka.add(v.x);
        kb.add(v.y);
    }
    a = warp_sum(ka.s);
    b = warp_sum(kb.s);
    if (lane == 0) {
        sh_a[warp] = a;
        sh_b[warp] = b;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        KahanPair fa{0.0, 0.0}, fb{0.0, 0.0};
        for (int w = 0; w < nwarps; ++w) {
            fa.add(sh_a[w]);
            fb.add(sh_b[w]);
        }
        out[2 * point] = fa.s;
        out[2 * point + 1] = fb.s;
        tickets[point] = 0u;               // re-arm for the next launch
    }
}

}  // namespace nmchb
