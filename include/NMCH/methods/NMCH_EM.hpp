// NMCH/methods/NMCH_EM.hpp -- "exact method" family, source-compatible with the reference
// (/root/reference/include/NMCH/methods/NMCH_EM.hpp:19-128).  All variants map to the one EM engine.
#ifndef NMCH_EXACT_METHOD_HPP
#define NMCH_EXACT_METHOD_HPP

#include "NMCH/methods/NMCH.hpp"
#include "NMCH/utils/utils.hpp"

namespace nmch::methods {

template <typename rnd_state>
class NMCH_EM_K1 : public NMCH<rnd_state> {
public:
    NMCH_EM_K1(int NTPB, int NB, float T, float S_0, float v_0, float r, float k, float rho, float theta, float sigma, int N);
    virtual void finalize() override;
    virtual void print_stats() override;
    virtual ~NMCH_EM_K1() = default;

    float get_execution_time() const { return Tim_exec; }

    /* reference NMCH_EM.hpp:49-54 */
    float get_err() const
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
    void run_init(unsigned long long seed);
    void run_compute();
};

template <typename rnd_state>
class NMCH_EM_K1_MM : public NMCH_EM_K1<rnd_state> {
public:
    NMCH_EM_K1_MM(int NTPB, int NB, float T, float S_0, float v_0, float r, float k, float rho, float theta, float sigma, int N);
    virtual void compute() override;
    virtual void init(unsigned long long seed) override;
    virtual ~NMCH_EM_K1_MM() = default;
};

template <typename rnd_state>
class NMCH_EM_K2_MM : public NMCH_EM_K1_MM<rnd_state> {
public:
    NMCH_EM_K2_MM(int NTPB, int NB, float T, float S_0, float v_0, float r, float k, float rho, float theta, float sigma, int N);
    virtual void compute() override;
    virtual ~NMCH_EM_K2_MM() = default;
};

template <typename rnd_state>
class NMCH_EM_K3_MM : public NMCH_EM_K2_MM<rnd_state> {
public:
    NMCH_EM_K3_MM(int NTPB, int NB, float T, float S_0, float v_0, float r, float k, float rho, float theta, float sigma, int N);
    virtual void compute() override;
    virtual ~NMCH_EM_K3_MM() = default;
};

}  // namespace nmch::methods

#endif  // NMCH_EXACT_METHOD_HPP
