/*
 * nmch_oracle.h -- CPU oracle for the Heston Monte-Carlo hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and only as the checker.
 *
 * It restates, in plain C, what the reference (edo01/NMCH) computes on the
 * GPU, with the third-party cuRAND device API (CUDA 12.9, cuRAND 10.3.10)
 * restated from its published algorithm.  Citations: reference paths are
 * relative to /root/reference, cuRAND paths to /usr/local/cuda/include.
 *
 * Parity pins: the reference ships no tests / golden vectors,
 * so the oracle is pinned against (i) the Random123 Philox4x32-10 KATs,
 * (ii) cuRAND's own headers compiled for the host (oracle/curand_host.cpp,
 * fixtures in tests/golden/), (iii) the reference's CUDA build run on a B200
 * (oracle/_ref/nmch_ref_harness, fixtures in tests/golden/ref_cuda_*.json).
 */
#ifndef NMCH_ORACLE_H
#define NMCH_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* RNG stream kinds */
enum { ORC_RNG_XORWOW = 0, ORC_RNG_PHILOX = 1, ORC_RNG_MRG32K3A = 2,
       /* NOT a cuRAND stream: the product's opt-in dense-draw mapping (three (23-bit, 19-bit) draws per Philox block,
        * nmch_b200/csrc/fe_kernels.cu), restated so that mode can be checked path by path; normal pairs only */
       ORC_RNG_PHILOX_DENSE = 3 };
/* variance floor g(.) : README.md:37-40 ; only abs is coded in the reference */
enum { ORC_FLOOR_ABS = 0, ORC_FLOOR_PLUS = 1 };

/* The 11 constructor values of nmch::methods::NMCH (include/NMCH/methods/NMCH.hpp:42) */
typedef struct {
    float T, S_0, v_0, r, k, rho, theta, sigma;
    int   N;
} orc_params_t;

/* Generator state; mirrors what cuRAND keeps (curand_kernel.h:150-156,
 * curand_philox4x32_x.h:93-102) so streams can continue across calls. */
typedef struct {
    int      kind;
    /* xorwow */
    uint32_t d, v[5];
    /* philox */
    uint32_t ctr[4], key[2], out[4];
    int      pos;
    /* mrg32k3a (curand_kernel.h:208-215) */
    uint32_t s1[3], s2[3];
    /* dense mapping: global step index */
    uint64_t dense_step;
    /* Box-Muller caches (curand_normal.h:313-326, 581-596) */
    int      bm_flag;
    float    bm_extra;
    int      bm_flag_d;
    double   bm_extra_d;
} orc_rng_t;

/* --- integer streams ----------------------------------------------------- */
void     orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
void     orc_rng_init(orc_rng_t *s, int kind, uint64_t seed, uint64_t subsequence, uint64_t offset);
uint32_t orc_rng_next(orc_rng_t *s);
/* --- float transforms ---------------------------------------------------- */
float    orc_uniform(orc_rng_t *s);
void     orc_normal2(orc_rng_t *s, float *gx, float *gy);
float    orc_normal(orc_rng_t *s);
double   orc_normal_double(orc_rng_t *s);
unsigned orc_poisson(orc_rng_t *s, double lambda);
float    orc_gamma(orc_rng_t *s, float alpha);

/* --- FE (src/NMCH/methods/NMCH_FE.cu:145-175) ----------------------------- */
/* Simulates paths [first_path, first_path+n_paths) for `calls` consecutive
 * compute() calls (streams continue, NMCH_FE.cu:303) and returns the terminal
 * S and V of the LAST call (arrays of n_paths, may be NULL) and its raw
 * payoff moments sum[(S-K)+], sum[((S-K)+)^2] accumulated in double. */
void orc_fe_run(const orc_params_t *p, int rng_kind, int floor_kind, uint64_t seed,
                uint64_t first_path, uint64_t n_paths, int calls,
                float *S_out, float *V_out, double *sum, double *sumsq, int threads);

/* same with curand_init's `offset` argument (u32 draws to skip; the reference always passes 0) */
void orc_fe_run_at(const orc_params_t *p, int rng_kind, int floor_kind, uint64_t seed, uint64_t offset,
                   uint64_t first_path, uint64_t n_paths, int calls,
                   float *S_out, float *V_out, double *sum, double *sumsq, int threads);

/* FE paths with the pathwise tangent dS_T/dv_0 (double) -- the checker of nmch_engine_compute_greeks; the step is
 * NMCH_FE.cu:156-163, its derivative is new (the reference has no sensitivities). */
void orc_fe_tangent_run(const orc_params_t *p, int rng_kind, int floor_kind, uint64_t seed,
                        uint64_t first_path, uint64_t n_paths, float *S_out, float *V_out, double *B_out, int threads);

/* The exploration sweep (src/NMCH/test/exploration.cu:71-88): per point set_k/theta/sigma + compute()
 * on continued streams.  sums has 2*n_points entries (raw sum, raw sum of squares per point). */
void orc_fe_sweep(const orc_params_t *p, int rng_kind, int floor_kind, uint64_t seed,
                  uint64_t first_path, uint64_t n_paths, int n_points, const float *k, const float *theta,
                  const float *sigma, double *sums, int threads);

/* --- EM (src/NMCH/methods/NMCH_EM.cu:11-55, 213-260) ---------------------- */
void orc_em_run(const orc_params_t *p, int rng_kind, uint64_t seed,
                uint64_t first_path, uint64_t n_paths, int calls,
                float *S_out, float *V_out, double *sum, double *sumsq, int threads);

/* Scheme-level restatement of the same EM recursion with exact samplers
 * (double-precision inversion Poisson, Marsaglia-Tsang gamma, own splitmix
 * stream).  NOT the reference's draws: used as the statistical oracle for the
 * native EM kernel (SURVEY.md §7 hard part 4). */
void orc_em_exact_run(const orc_params_t *p, uint64_t seed, uint64_t n_paths,
                      double *sum, double *sumsq, double *sum_ST, int threads);

/* --- QE-M, the product's large-step third method (no reference counterpart): restates the KERNEL's scheme and
 * draw mapping (nmch_b200/csrc/qe_kernels.cu) with libm, for per-path checks; `call` is the engine's call counter */
void orc_qe_run(const orc_params_t *p, uint64_t seed, uint64_t first_path, uint64_t n_paths, uint32_t call,
                float *S_out, float *V_out, double *sum, double *sumsq, int threads);

/* --- host statistics ------------------------------------------------------ */
/* include/NMCH/methods/NMCH_FE.hpp:50-55 (float/double mix reproduced) */
float  orc_get_err(int state_numbers, float strike_price, float price_squared);
/* src/NMCH/utils/utils.cu:5-25 and the "true price" of NMCH_FE.cu:336-338 */
double orc_NP(double x);
float  orc_print_true_price(float S_0, float K, float r, float sigma);
/* Semi-analytic Heston European call (characteristic function, "little trap") */
double orc_heston_call(double S0, double K, double v0, double r, double kappa,
                       double theta, double sigma, double rho, double T);

/* --- exploration grid (src/NMCH/test/exploration.cu:46-52, 71-88) ---------- */
/* Enumerates the reference's float-accumulated sweep in launch order
 * (sigma outer, theta, k inner), applying the 20*k*theta < sigma^2 skip.
 * Returns the number of points written (<= cap). */
int orc_exploration_grid(int steps, int apply_filter, float *k, float *theta, float *sigma, int cap);

void orc_em_native_run(const orc_params_t *p, uint64_t seed, uint64_t first_path, uint64_t n_paths, uint32_t call,
                       float *S_out, float *V_out, double *sum, double *sumsq, int threads);
uint64_t orc_rng_init_only(int rng_kind, uint64_t seed, uint64_t first_path, uint64_t n_paths, int threads);
int orc_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
