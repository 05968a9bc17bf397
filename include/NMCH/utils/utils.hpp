// NMCH/utils/utils.hpp -- same helpers as the reference (include/NMCH/utils/utils.hpp:10-21).
#ifndef NMCH_UTILS_HPP
#define NMCH_UTILS_HPP

#include <cmath>
#include <cstdio>

namespace nmch::utils {

// Normal CDF, Abramowitz-Stegun 26.2.17, as printed by print_stats (src/NMCH/utils/utils.cu:5-25)
double NP(double x);

// Semi-analytic Heston European call (new: the reference prints a Black-Scholes value as "true price")
double heston_call(double S0, double K, double v0, double r, double kappa, double theta, double sigma, double rho,
                   double T);

}  // namespace nmch::utils

namespace nmch::utils::cuda {

// Error convention of the reference (src/NMCH/utils/utils.cu:30-35): print file/line and exit(EXIT_FAILURE).
// `status` is an nmch_status from the C ABI (0 = ok); the engine's own detail string is printed as well.
void checkCUDA(int status, const char *file, int line);

}  // namespace nmch::utils::cuda

#endif  // NMCH_UTILS_HPP
