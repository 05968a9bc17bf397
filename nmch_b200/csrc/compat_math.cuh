// compat_math.cuh -- the reference's FP32 arithmetic, pinned: shared by every draw-compatible (validation) kernel.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "kernels.cuh"

namespace nmchb {

// cuRAND's Box-Muller (curand_normal.h:70-87): IEEE logf/sqrtf, fast __sincosf, x pairs with sin.
__device__ __forceinline__ void box_muller_compat(uint32_t x, uint32_t y, float &gx, float &gy)
{
    const float u = __fmaf_rn((float)x, 2.3283064e-10f, 2.3283064e-10f / 2.0f);
    const float v = __fmaf_rn((float)y, 2.3283064e-10f * 6.2831855f, (2.3283064e-10f * 6.2831855f) / 2.0f);
    const float s = sqrtf(__fmul_rn(-2.0f, logf(u)));
    float sn, cs;
    __sincosf(v, &sn, &cs);
    gx = __fmul_rn(sn, s);
    gy = __fmul_rn(cs, s);
}

// The reference's update (NMCH_FE.cu:160-162) with the FMA contraction nvcc applies to that source text
// (SURVEY.md Appendix B), pinned with explicit intrinsics so it cannot drift.
template <int FLOOR>
__device__ __forceinline__ void fe_step_compat(float &S, float &V, float gx, float gy, float r, float k,
                                               float rho, float theta, float sigma, float dt, float sqrt_dt,
                                               float sqrt_rho)
{
    const float sv = __fsqrt_rn(V);
    float a = __fmul_rn(r, S);
    a = __fmaf_rn(a, dt, S);
    float z = __fmul_rn(gy, sqrt_rho);
    z = __fmaf_rn(gx, rho, z);
    float b = __fmul_rn(sv, S);
    b = __fmul_rn(b, sqrt_dt);
    const float Sn = __fmaf_rn(b, z, a);
    float c = __fsub_rn(theta, V);
    c = __fmul_rn(c, k);
    c = __fmaf_rn(c, dt, V);
    float e = __fmul_rn(sv, sigma);
    e = __fmul_rn(e, sqrt_dt);
    const float Vn = __fmaf_rn(gx, e, c);
    S = Sn;
    V = (FLOOR == kFloorAbs) ? fabsf(Vn) : fmaxf(Vn, 0.0f);
}

}  // namespace nmchb
