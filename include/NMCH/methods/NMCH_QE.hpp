// NMCH/methods/NMCH_QE.hpp -- third method family (no reference counterpart): Andersen's
// quadratic-exponential scheme with martingale correction.  Same lifecycle and getters as the FE / EM families, so it
// slots into the same user code; meant for LARGE steps (N = 50..100 reaches the accuracy of FE at N = 1000).
// Only the native Philox tag (curandStatePhilox4_32_10_t) is meaningful: there is no reference stream to match.
#ifndef NMCH_QUADRATIC_EXPONENTIAL_HPP
#define NMCH_QUADRATIC_EXPONENTIAL_HPP

#include "NMCH/methods/NMCH.hpp"
#include "NMCH/utils/utils.hpp"

namespace nmch::methods {

template <typename rnd_state>
class NMCH_QE_K1_MM : public NMCH<rnd_state> {
public:
    NMCH_QE_K1_MM(int NTPB, int NB, float T, float S_0, float v_0, float r, float k, float rho, float theta, float sigma, int N);
    virtual void compute() override;
    virtual void init(unsigned long long seed) override;
    virtual void finalize() override;
    virtual void print_stats() override;
    virtual ~NMCH_QE_K1_MM() = default;

    float get_execution_time() const { return Tim_exec; }
    float get_err() const      /* the same 95% half-width formula as the FE / EM families (reference NMCH_FE.hpp:50-55) */
    {
        float err = 1.96 * sqrt((double)(1.0f / (this->state_numbers - 1)) *
                                (this->state_numbers * this->price_squared - (this->strike_price * this->strike_price))) /
                    sqrt((double)this->state_numbers);
        return err;
    }

protected:
    int state_numbers;
    float Tim_exec;
    float Tim_init;
};

}  // namespace nmch::methods

#endif  // NMCH_QUADRATIC_EXPONENTIAL_HPP
