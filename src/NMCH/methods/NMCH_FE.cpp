// NMCH_FE.cpp -- forward-Euler method family over the engine (reference: src/NMCH/methods/NMCH_FE.cu:312-689,
// the host halves; its kernels :6-307 are replaced by nmch_b200/csrc/fe_kernels.cu).
#include "NMCH/methods/NMCH_FE.hpp"

namespace nmch::methods {

// ---- shared behaviour of every FE class ----------------------------------------------------------------
template <typename S>
NMCH_FE_K1<S>::NMCH_FE_K1(int NTPB, int NB, float T, float S_0, float v_0, float r, float k, float rho, float theta,
                          float sigma, int N)
    : NMCH<S>(NTPB, NB, T, S_0, v_0, r, k, rho, theta, sigma, N), Tim_exec(0.0f), Tim_init(0.0f)
{
    state_numbers = NTPB * NB;
}

template <typename S> void NMCH_FE_K1<S>::run_init(unsigned long long seed) { this->engine_init(NMCH_METHOD_FE, seed, &Tim_init); }
template <typename S> void NMCH_FE_K1<S>::run_compute() { this->engine_compute(&Tim_exec); }
template <typename S> void NMCH_FE_K1<S>::finalize() { this->engine_finalize(); }

template <typename S>
void NMCH_FE_K1<S>::print_stats()
{
    // the reference prints a Black-Scholes value (vol := sigma, T := 1) as "true price" (NMCH_FE.cu:336-338)
    float real_price = this->S_0 * nmch::utils::NP((this->r + 0.5 * this->sigma * this->sigma) / this->sigma) -
                       this->K * expf(-this->r) * nmch::utils::NP((this->r - 0.5 * this->sigma * this->sigma) / this->sigma);
    NMCH<S>::print_stats();
    printf("METHOD: FORWARD-EULER\n");
    printf("The estimated price E[X] is equal to %f\n", this->strike_price);
    printf("The estimated E[X^2] is equal to %f\n", this->price_squared);
    printf("The true price %f\n", real_price);
    printf("error associated to a confidence interval of 95%% = %f\n", get_err());
    printf("Execution time %f ms\n", Tim_exec);
    printf("Initialization time %f ms\n", Tim_init);
}

// ---- the named variants: identical engine underneath ---------------------------------------------------
#define NMCH_FE_CTOR(CLASS, BASE)                                                                                    \
    template <typename S>                                                                                            \
    CLASS<S>::CLASS(int NTPB, int NB, float T, float S_0, float v_0, float r, float k, float rho, float theta,        \
                    float sigma, int N)                                                                              \
        : BASE<S>(NTPB, NB, T, S_0, v_0, r, k, rho, theta, sigma, N)                                                 \
    {                                                                                                                \
    }

NMCH_FE_CTOR(NMCH_FE_K1_MM, NMCH_FE_K1)
NMCH_FE_CTOR(NMCH_FE_K2_MM, NMCH_FE_K1_MM)
NMCH_FE_CTOR(NMCH_FE_K3_MM, NMCH_FE_K2_MM)
NMCH_FE_CTOR(NMCH_FE_K1_PgM, NMCH_FE_K1)
NMCH_FE_CTOR(NMCH_FE_K1_PiM, NMCH_FE_K1)

template <typename S> void NMCH_FE_K1_MM<S>::init(unsigned long long seed) { this->run_init(seed); }
// the three classes below launch FE_k1 in the reference (NMCH_FE.cu:516-546, 556-612, 622-689): the K1 moment quirk is theirs
template <typename S> void NMCH_FE_K1_MM<S>::compute() { this->run_compute(); this->apply_legacy_k1_moment(); }
template <typename S> void NMCH_FE_K2_MM<S>::compute() { this->run_compute(); }
template <typename S> void NMCH_FE_K3_MM<S>::compute() { this->run_compute(); }
template <typename S> void NMCH_FE_K1_PgM<S>::init(unsigned long long seed) { this->run_init(seed); }
template <typename S> void NMCH_FE_K1_PgM<S>::compute() { this->run_compute(); this->apply_legacy_k1_moment(); }
template <typename S> void NMCH_FE_K1_PiM<S>::init(unsigned long long seed) { this->run_init(seed); }
template <typename S> void NMCH_FE_K1_PiM<S>::compute() { this->run_compute(); this->apply_legacy_k1_moment(); }
template <typename S> void NMCH_FE_K1_PiM<S>::finalize() { NMCH_FE_K1<S>::finalize(); }

NMCH_FE_K2_PHILOX_MM::NMCH_FE_K2_PHILOX_MM(int NTPB, int NB, float T, float S_0, float v_0, float r, float k, float rho,
                                           float theta, float sigma, int N)
    : NMCH_FE_K1_MM<curandStatePhilox4_32_10_t>(NTPB, NB, T, S_0, v_0, r, k, rho, theta, sigma, N)
{
}
void NMCH_FE_K2_PHILOX_MM::compute() { this->run_compute(); }

// explicit instantiations, the same tag set as the reference (NMCH_FE.cu:352-354 and following)
#define NMCH_FE_INSTANTIATE(TAG)            \
    template class NMCH_FE_K1<TAG>;         \
    template class NMCH_FE_K1_MM<TAG>;      \
    template class NMCH_FE_K2_MM<TAG>;      \
    template class NMCH_FE_K3_MM<TAG>;      \
    template class NMCH_FE_K1_PgM<TAG>;     \
    template class NMCH_FE_K1_PiM<TAG>;

NMCH_FE_INSTANTIATE(curandStateXORWOW_t)
NMCH_FE_INSTANTIATE(curandStateMRG32k3a_t)
NMCH_FE_INSTANTIATE(curandStatePhilox4_32_10_t)

}  // namespace nmch::methods
