/*
 * curand_host.cpp -- the third-party cuRAND device API (CUDA 12.9, cuRAND
 * 10.3.10; the only dependency of the reference's hot path, see
 * /root/reference/src/NMCH/random/random.cu:9 and the call sites in
 * NMCH_FE.cu:158 / NMCH_EM.cu:237,260) compiled FOR THE HOST.
 *
 * TEST INFRASTRUCTURE ONLY.  The header-only API is host-compilable because
 * its QUALIFIERS macro is overridable (curand_kernel.h:60-62) and it carries
 * host copies of the XORWOW skip-ahead tables (curand_precalc.h:935).  The
 * integer streams it produces here are bit-identical to the device's; the
 * float transforms take the header's host branches (sinf/cosf, expf/logf).
 *
 * Used by tests/golden/make_golden.py to pin oracle/nmch_oracle.c (our own
 * restatement) and committed as fixtures; never linked into the product.
 */
#include <cuda_runtime.h>
#define QUALIFIERS static inline __host__ __device__
#include <curand_kernel.h>

#include <cstdint>

extern "C" {

/* raw XORWOW state after curand_init(seed, subsequence, offset) */
void crh_xorwow_init(unsigned long long seed, unsigned long long subseq, unsigned long long offset,
                     uint32_t *d, uint32_t v[5])
{
    curandStateXORWOW_t s;
    curand_init(seed, subseq, offset, &s);
    *d = s.d;
    for (int i = 0; i < 5; ++i) v[i] = s.v[i];
}

/* first n 32-bit draws of curand() */
void crh_u32(int kind, unsigned long long seed, unsigned long long subseq, unsigned long long offset,
             int n, uint32_t *out)
{
    if (kind == 0) {
        curandStateXORWOW_t s;
        curand_init(seed, subseq, offset, &s);
        for (int i = 0; i < n; ++i) out[i] = curand(&s);
    } else if (kind == 1) {
        curandStatePhilox4_32_10_t s;
        curand_init(seed, subseq, offset, &s);
        for (int i = 0; i < n; ++i) out[i] = curand(&s);
    } else {
        curandStateMRG32k3a_t s;
        curand_init(seed, subseq, offset, &s);
        for (int i = 0; i < n; ++i) out[i] = curand(&s);
    }
}

/* raw MRG32k3a state after curand_init */
void crh_mrg_init(unsigned long long seed, unsigned long long subseq, unsigned long long offset, uint32_t st[6])
{
    curandStateMRG32k3a_t s;
    curand_init(seed, subseq, offset, &s);
    for (int i = 0; i < 3; ++i) { st[i] = s.s1[i]; st[3 + i] = s.s2[i]; }
}

/* XORWOW raw state after n curand_normal2 calls (pins skip + stream position) */
void crh_xorwow_state_after_normal2(unsigned long long seed, unsigned long long subseq, int n,
                                    uint32_t *d, uint32_t v[5])
{
    curandStateXORWOW_t s;
    curand_init(seed, subseq, 0, &s);
    for (int i = 0; i < n; ++i) (void)curand_normal2(&s);
    *d = s.d;
    for (int i = 0; i < 5; ++i) v[i] = s.v[i];
}

/* n pairs from curand_normal2 */
void crh_normal2(int kind, unsigned long long seed, unsigned long long subseq, int n, float *out)
{
    if (kind == 0) {
        curandStateXORWOW_t s;
        curand_init(seed, subseq, 0, &s);
        for (int i = 0; i < n; ++i) { float2 g = curand_normal2(&s); out[2 * i] = g.x; out[2 * i + 1] = g.y; }
    } else if (kind == 1) {
        curandStatePhilox4_32_10_t s;
        curand_init(seed, subseq, 0, &s);
        for (int i = 0; i < n; ++i) { float2 g = curand_normal2(&s); out[2 * i] = g.x; out[2 * i + 1] = g.y; }
    } else {
        curandStateMRG32k3a_t s;
        curand_init(seed, subseq, 0, &s);
        for (int i = 0; i < n; ++i) { float2 g = curand_normal2(&s); out[2 * i] = g.x; out[2 * i + 1] = g.y; }
    }
}

}  // extern "C"
/* n draws of curand_poisson(lambda) from one stream */
template <typename S>
static void poisson_t(unsigned long long seed, unsigned long long subseq, double lambda, int n, unsigned *out)
{
    S s;
    curand_init(seed, subseq, 0, &s);
    for (int i = 0; i < n; ++i) out[i] = curand_poisson(&s, lambda);
}
extern "C" void crh_poisson(int kind, unsigned long long seed, unsigned long long subseq, double lambda, int n, unsigned *out)
{
    if (kind == 0)      poisson_t<curandStateXORWOW_t>(seed, subseq, lambda, n, out);
    else if (kind == 1) poisson_t<curandStatePhilox4_32_10_t>(seed, subseq, lambda, n, out);
    else                poisson_t<curandStateMRG32k3a_t>(seed, subseq, lambda, n, out);
}

/* mixed sequence exercising the caches: uniform, normal, normal_double, normal, uniform ... */
template <typename S>
static void mixed_t(unsigned long long seed, unsigned long long subseq, int n, double *out)
{
    S s;
    curand_init(seed, subseq, 0, &s);
    for (int i = 0; i < n; ++i) {
        switch (i % 5) {
        case 0: out[i] = curand_uniform(&s); break;
        case 1: out[i] = curand_normal(&s); break;
        case 2: out[i] = curand_normal_double(&s); break;
        case 3: out[i] = curand_normal(&s); break;
        default: out[i] = (double)curand(&s); break;
        }
    }
}
extern "C" void crh_mixed(int kind, unsigned long long seed, unsigned long long subseq, int n, double *out)
{
    if (kind == 0)      mixed_t<curandStateXORWOW_t>(seed, subseq, n, out);
    else if (kind == 1) mixed_t<curandStatePhilox4_32_10_t>(seed, subseq, n, out);
    else                mixed_t<curandStateMRG32k3a_t>(seed, subseq, n, out);
}
