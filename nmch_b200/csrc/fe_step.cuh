// fe_step.cuh -- the native FE step, shared by the path kernels (fe_kernels.cu) and the tangent kernel
// (greeks_kernels.cu), so that both advance (S, V) with the same instructions on the same Philox words.
#pragma once
#include "kernels.cuh"

namespace nmchb {

// PRECISE_V: V' = g(V - kdt*V + (vb + q sin * vs)) with the product rounded once, instead of the folded
// V*va + vb + ... -- one more FP32 operation.  The folded va = 1 - k*dt carries a rounding error of up to 3e-8, i.e. a
// relative error of up to 3e-8 / (k dt) in the mean-reversion speed; fold_fe_point compensates vb so that the
// long-run level stays theta, which is all the 3-standard-error modes need.  The XORWOW_FAST mode promises 1e-5
// against the reference on identical draws and takes the exact form.
template <int FLOOR, bool PRECISE_V = false>
__device__ __forceinline__ void fe_step_native(float &S, float &V, uint32_t wa, uint32_t wb, float rdt,
                                               float zr, float zc, const FePoint &pc)
{
    // uniforms as cuRAND forms them (curand_uniform.h:69-72, curand_normal.h:72-75): u = x 2^-32 + 2^-33 in (0, 1],
    // angle = y (2 pi 2^-32) -- an integer-to-float conversion (I2FP) and one FFMA / FMUL.  Against bit
    // splicing ((w >> 9) | 0x3f800000, then a subtraction) this costs the same instruction count, but I2FP issues beside
    // the FP32 work where LEA.HI does not (profiles/r02_pipe_rates2.txt: -1 % on the kernel), and the native mode now
    // feeds the same uniforms as the draw-compatible modes into its fast transforms.
    constexpr float k2Pow32Inv = 2.3283064e-10f, k2Pow32Inv2Pi = 2.3283064e-10f * 6.2831855f;
    const float u = fmaf(__uint2float_rn(wa), k2Pow32Inv, k2Pow32Inv * 0.5f);
    const float l2 = lg2_approx(u);                   // <= 0
    const float q = sqrt_approx(-(V * l2));           // sqrt(V) * sqrt(-lg2 u)
    const float ang = __uint2float_rn(wb) * k2Pow32Inv2Pi;   // cuRAND adds half a step (7e-10 rad, below the angle's ulp)
    const float gs = q * sin_approx(ang);
    const float gc = q * cos_approx(ang);
    float m = fmaf(gs, zr, rdt);                      // relative increment; S' = S + S*m keeps r*dt at full precision
    m = fmaf(gc, zc, m);                              // (a folded 1 + r*dt would round by up to 6e-8 every step, a
    S = fmaf(S, m, S);                                //  systematic drift of N * 6e-8 on S_T)
    float vn;
    if constexpr (PRECISE_V) {
        vn = fmaf(-pc.kdt, V, V) + fmaf(gs, pc.vs, pc.vb);
    } else {
        vn = fmaf(V, pc.va, pc.vb);
        vn = fmaf(gs, pc.vs, vn);
    }
    V = (FLOOR == kFloorAbs) ? fabsf(vn) : fmaxf(vn, 0.0f);
}

// The same step carrying the pathwise tangent in v_0:  A = dV/dv_0,  B = dS/dv_0  (A_0 = 1, B_0 = 0).
// (S, V) take EXACTLY the instructions of fe_step_native (same expressions in the same order), so a tangent pass
// reproduces the path pass bit for bit.  Differentiating the step:  q = sqrt(V) e  =>  dq = q/(2V) dV, hence with
// h = A / (2V):   d(gs) = gs h,  d(gc) = gc h,   dm = (gs zr + gc zc) h,
//   B' = B (1 + m) + S dm,      A'' = A va + gs vs h,      A' = g'(vn) A''   (g' = sign for |.|, 1{vn > 0} for (.)+).
// V = 0 (the (.)+ floor parks paths there) has q = 0 and an infinite dq/dV; A is then 0 already (the floor's derivative
// zeroed it in the step that parked the path), and the 0 * inf is taken as 0.
template <int FLOOR>
__device__ __forceinline__ void fe_step_native_tangent(float &S, float &V, float &A, float &B, uint32_t wa, uint32_t wb,
                                                       float rdt, float zr, float zc, const FePoint &pc)
{
    constexpr float k2Pow32Inv = 2.3283064e-10f, k2Pow32Inv2Pi = 2.3283064e-10f * 6.2831855f;
    const float u = fmaf(__uint2float_rn(wa), k2Pow32Inv, k2Pow32Inv * 0.5f);
    const float l2 = lg2_approx(u);
    const float q = sqrt_approx(-(V * l2));
    const float ang = __uint2float_rn(wb) * k2Pow32Inv2Pi;
    const float gs = q * sin_approx(ang);
    const float gc = q * cos_approx(ang);
    float m = fmaf(gs, zr, rdt);
    m = fmaf(gc, zc, m);
    const float h = (V > 0.0f) ? __fdividef(0.5f * A, V) : 0.0f;
    const float dm = fmaf(gc, zc, gs * zr) * h;
    B = fmaf(S, dm, fmaf(B, m, B));                   // uses S before its update
    S = fmaf(S, m, S);
    float vn = fmaf(V, pc.va, pc.vb);
    vn = fmaf(gs, pc.vs, vn);
    const float an = fmaf(gs * pc.vs, h, A * pc.va);
    if (FLOOR == kFloorAbs) {
        A = (vn < 0.0f) ? -an : an;
        V = fabsf(vn);
    } else {
        A = (vn > 0.0f) ? an : 0.0f;
        V = fmaxf(vn, 0.0f);
    }
}

}  // namespace nmchb
