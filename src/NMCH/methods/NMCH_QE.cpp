// NMCH_QE.cpp -- the QE-M method family over the engine (kernel: nmch_b200/csrc/qe_kernels.cu).
#include "NMCH/methods/NMCH_QE.hpp"

namespace nmch::methods {

template <typename S>
NMCH_QE_K1_MM<S>::NMCH_QE_K1_MM(int NTPB, int NB, float T, float S_0, float v_0, float r, float k, float rho, float theta,
                                float sigma, int N)
    : NMCH<S>(NTPB, NB, T, S_0, v_0, r, k, rho, theta, sigma, N), Tim_exec(0.0f), Tim_init(0.0f)
{
    state_numbers = NTPB * NB;
}

template <typename S> void NMCH_QE_K1_MM<S>::init(unsigned long long seed) { this->engine_init(NMCH_METHOD_QE, seed, &Tim_init); }
template <typename S> void NMCH_QE_K1_MM<S>::compute() { this->engine_compute(&Tim_exec); }
template <typename S> void NMCH_QE_K1_MM<S>::finalize() { this->engine_finalize(); }

template <typename S>
void NMCH_QE_K1_MM<S>::print_stats()
{
    NMCH<S>::print_stats();
    printf("METHOD: QUADRATIC-EXPONENTIAL\n");
    printf("The estimated price E[X] is equal to %f\n", this->strike_price);
    printf("The estimated E[X^2] is equal to %f\n", this->price_squared);
    printf("The true price %f\n", (float)nmch::utils::heston_call(this->S_0, this->K, this->v_0, this->r, this->k, this->theta,
                                                                 this->sigma, this->rho, this->T) * expf(this->r * this->T));
    printf("error associated to a confidence interval of 95%% = %f\n", get_err());
    printf("Execution time %f ms\n", Tim_exec);
    printf("Initialization time %f ms\n", Tim_init);
}

template class NMCH_QE_K1_MM<curandStateXORWOW_t>;
template class NMCH_QE_K1_MM<curandStateMRG32k3a_t>;
template class NMCH_QE_K1_MM<curandStatePhilox4_32_10_t>;

}  // namespace nmch::methods
