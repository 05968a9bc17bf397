// NMCH_EM.cpp -- "exact method" family over the engine (reference: src/NMCH/methods/NMCH_EM.cu:376-575, host halves).
#include "NMCH/methods/NMCH_EM.hpp"

namespace nmch::methods {

template <typename S>
NMCH_EM_K1<S>::NMCH_EM_K1(int NTPB, int NB, float T, float S_0, float v_0, float r, float k, float rho, float theta,
                          float sigma, int N)
    : NMCH<S>(NTPB, NB, T, S_0, v_0, r, k, rho, theta, sigma, N), Tim_exec(0.0f), Tim_init(0.0f)
{
    state_numbers = NTPB * NB;
}

template <typename S> void NMCH_EM_K1<S>::run_init(unsigned long long seed) { this->engine_init(NMCH_METHOD_EM, seed, &Tim_init); }
template <typename S> void NMCH_EM_K1<S>::run_compute() { this->engine_compute(&Tim_exec); }
template <typename S> void NMCH_EM_K1<S>::finalize() { this->engine_finalize(); }

template <typename S>
void NMCH_EM_K1<S>::print_stats()
{
    float real_price = this->S_0 * nmch::utils::NP((this->r + 0.5 * this->sigma * this->sigma) / this->sigma) -
                       this->K * expf(-this->r) * nmch::utils::NP((this->r - 0.5 * this->sigma * this->sigma) / this->sigma);
    NMCH<S>::print_stats();
    printf("METHOD: EXACT-METHOD\n");
    printf("The estimated price E[X] is equal to %f\n", this->strike_price);
    printf("The estimated E[X^2] is equal to %f\n", this->price_squared);
    printf("The true price %f\n", real_price);
    printf("error associated to a confidence interval of 95%% = %f\n", get_err());
    printf("Execution time %f ms\n", Tim_exec);
    printf("Initialization time %f ms\n", Tim_init);
}

#define NMCH_EM_CTOR(CLASS, BASE)                                                                                    \
    template <typename S>                                                                                            \
    CLASS<S>::CLASS(int NTPB, int NB, float T, float S_0, float v_0, float r, float k, float rho, float theta,        \
                    float sigma, int N)                                                                              \
        : BASE<S>(NTPB, NB, T, S_0, v_0, r, k, rho, theta, sigma, N)                                                 \
    {                                                                                                                \
    }

NMCH_EM_CTOR(NMCH_EM_K1_MM, NMCH_EM_K1)
NMCH_EM_CTOR(NMCH_EM_K2_MM, NMCH_EM_K1_MM)
NMCH_EM_CTOR(NMCH_EM_K3_MM, NMCH_EM_K2_MM)

template <typename S> void NMCH_EM_K1_MM<S>::init(unsigned long long seed) { this->run_init(seed); }
template <typename S> void NMCH_EM_K1_MM<S>::compute() { this->run_compute(); this->apply_legacy_k1_moment(); }   // EM_k1, NMCH_EM.cu:129-131
template <typename S> void NMCH_EM_K2_MM<S>::compute() { this->run_compute(); }
template <typename S> void NMCH_EM_K3_MM<S>::compute() { this->run_compute(); }

#define NMCH_EM_INSTANTIATE(TAG)        \
    template class NMCH_EM_K1<TAG>;     \
    template class NMCH_EM_K1_MM<TAG>;  \
    template class NMCH_EM_K2_MM<TAG>;  \
    template class NMCH_EM_K3_MM<TAG>;

NMCH_EM_INSTANTIATE(curandStateXORWOW_t)
NMCH_EM_INSTANTIATE(curandStateMRG32k3a_t)
NMCH_EM_INSTANTIATE(curandStatePhilox4_32_10_t)

}  // namespace nmch::methods
