// NMCH/methods/NMCH_FE.hpp -- forward-Euler method family, source-compatible with the reference
// (/root/reference/include/NMCH/methods/NMCH_FE.hpp:20-189).  The reference's variants differ in memory space
// (managed / pageable / pinned), reduction (shared-memory tree vs warp shuffle) and where the cuRAND state
// lives; here every name maps to the one fused sm_100a kernel, whose result does not depend on those choices.
// One documented difference: the reference's K1 classes store E[X^2]/n^2 in price_squared by accident
// (NMCH_FE.cu:56-58); by default all classes here store E[X^2], like its K2/K3 classes, and
// set_legacy_k1_moment(true) makes the K1 classes reproduce the reference's value exactly.
#ifndef NMCH_FW_EULER_HPP
#define NMCH_FW_EULER_HPP

#include "NMCH/methods/NMCH.hpp"
#include "NMCH/utils/utils.hpp"

namespace nmch::methods {

template <typename rnd_state>
class NMCH_FE_K1 : public NMCH<rnd_state> {
public:
    NMCH_FE_K1(int NTPB, int NB, float T, float S_0, float v_0, float r, float k, float rho, float theta, float sigma, int N);
    virtual void finalize() override;
    virtual void print_stats() override;
    virtual ~NMCH_FE_K1() = default;

    float get_execution_time() const { return Tim_exec; }

    /* The reference's 95% half-width, float/double mix included (NMCH_FE.hpp:50-55):
       1.96 * sqrt((1/(n-1)) * (n*E[X^2] - E[X]^2)) / sqrt(n) */
    float get_err() const
    {
        float err = 1.96 * sqrt((double)(1.0f / (this->state_numbers - 1)) *
                                (this->state_numbers * this->price_squared - (this->strike_price * this->strike_price))) /
                    sqrt((double)this->state_numbers);
        return err;
    }

protected:
    int state_numbers;   /* number of paths (one generator stream each) */
    float Tim_exec;      /* kernel + sync, ms */
    float Tim_init;      /* allocation + generator-state construction, ms */
    void run_init(unsigned long long seed);
    void run_compute();
};

template <typename rnd_state>
class NMCH_FE_K1_MM : public NMCH_FE_K1<rnd_state> {
public:
    NMCH_FE_K1_MM(int NTPB, int NB, float T, float S_0, float v_0, float r, float k, float rho, float theta, float sigma, int N);
    virtual void compute() override;
    virtual void init(unsigned long long seed) override;
    virtual ~NMCH_FE_K1_MM() = default;
};

template <typename rnd_state>
class NMCH_FE_K2_MM : public NMCH_FE_K1_MM<rnd_state> {
public:
    NMCH_FE_K2_MM(int NTPB, int NB, float T, float S_0, float v_0, float r, float k, float rho, float theta, float sigma, int N);
    virtual void compute() override;
    virtual ~NMCH_FE_K2_MM() = default;
};

/* Non-template in the reference (NMCH_FE.hpp:142): always the Philox generator. */
class NMCH_FE_K2_PHILOX_MM : public NMCH_FE_K1_MM<curandStatePhilox4_32_10_t> {
public:
    NMCH_FE_K2_PHILOX_MM(int NTPB, int NB, float T, float S_0, float v_0, float r, float k, float rho, float theta, float sigma, int N);
    virtual void compute() override;
    virtual ~NMCH_FE_K2_PHILOX_MM() = default;
};

template <typename rnd_state>
class NMCH_FE_K3_MM : public NMCH_FE_K2_MM<rnd_state> {
public:
    NMCH_FE_K3_MM(int NTPB, int NB, float T, float S_0, float v_0, float r, float k, float rho, float theta, float sigma, int N);
    virtual void compute() override;
    virtual ~NMCH_FE_K3_MM() = default;
};

template <typename rnd_state>
class NMCH_FE_K1_PgM : public NMCH_FE_K1<rnd_state> {
public:
    NMCH_FE_K1_PgM(int NTPB, int NB, float T, float S_0, float v_0, float r, float k, float rho, float theta, float sigma, int N);
    virtual void compute() override;
    virtual void init(unsigned long long seed) override;
    virtual ~NMCH_FE_K1_PgM() = default;
};

template <typename rnd_state>
class NMCH_FE_K1_PiM : public NMCH_FE_K1<rnd_state> {
public:
    NMCH_FE_K1_PiM(int NTPB, int NB, float T, float S_0, float v_0, float r, float k, float rho, float theta, float sigma, int N);
    virtual void compute() override;
    virtual void init(unsigned long long seed) override;
    virtual void finalize() override;
    virtual ~NMCH_FE_K1_PiM() = default;
};

}  // namespace nmch::methods

#endif  // NMCH_FW_EULER_HPP
