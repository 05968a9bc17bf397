// utils.cpp -- host helpers of the method API.
#include "NMCH/utils/utils.hpp"

#include <complex>
#include <cstdlib>

#include "nmch_b200.h"

namespace nmch::utils {

double NP(double x)
{
    // Abramowitz-Stegun 26.2.17 with the reference's constants (src/NMCH/utils/utils.cu:5-25), so the printed
    // "true price" line is digit-for-digit the reference's
    const double p = 0.2316419;
    const double b[5] = {0.319381530, -0.356563782, 1.781477937, -1.821255978, 1.330274429};
    const double inv_sqrt_2pi = 0.39894228;
    const double ax = x >= 0.0 ? x : -x;
    const double t = 1.0 / (1.0 + p * ax);
    const double poly = t * (t * (t * (t * (t * b[4] + b[3]) + b[2]) + b[1]) + b[0]);
    const double tail = inv_sqrt_2pi * std::exp(-x * x / 2.0) * poly;
    return x >= 0.0 ? 1.0 - tail : tail;
}

// Heston (1993) P1/P2 integrals with the Albrecher et al. branch-stable characteristic function,
// composite Gauss-Legendre on [0, 400].  Used for the exploration CSV's bias column and for reports.
double heston_call(double S0, double K, double v0, double r, double kappa, double theta, double sigma, double rho, double T)
{
    using cd = std::complex<double>;
    static const double gx[8] = {0.0950125098376374, 0.2816035507792589, 0.4580167776572274, 0.6178762444026438,
                                 0.7554044083550030, 0.8656312023878318, 0.9445750230732326, 0.9894009349916499};
    static const double gw[8] = {0.1894506104550685, 0.1826034150449236, 0.1691565193950025, 0.1495959888165767,
                                 0.1246289712555339, 0.0951585116824928, 0.0622535239386479, 0.0271524594117541};
    const double lnS0 = std::log(S0), lnK = std::log(K), s2 = sigma * sigma;
    auto integrand = [&](double phi, int j) {
        const double u = (j == 1) ? 0.5 : -0.5;
        const double b = (j == 1) ? kappa - rho * sigma : kappa;
        const cd iphi(0.0, phi);
        const cd rsi = rho * sigma * iphi;
        const cd d = std::sqrt((rsi - b) * (rsi - b) - s2 * (2.0 * u * iphi - phi * phi));
        const cd g = (b - rsi - d) / (b - rsi + d);
        const cd edT = std::exp(-d * T);
        const cd C = r * iphi * T + kappa * theta / s2 * ((b - rsi - d) * T - 2.0 * std::log((1.0 - g * edT) / (1.0 - g)));
        const cd D = (b - rsi - d) / s2 * (1.0 - edT) / (1.0 - g * edT);
        const cd f = std::exp(C + D * v0 + iphi * lnS0);
        return (std::exp(-iphi * lnK) * f / iphi).real();
    };
    // Panels of width 1/2 with 16 Gauss-Legendre nodes; the integrand decays exponentially, so stop once three
    // consecutive panels contribute less than 1e-16 (at most 400 / 0.5 panels, as the fixed rule used before).
    const double h = 0.5;
    const int max_panels = 800;
    double I1 = 0.0, I2 = 0.0;
    int quiet = 0;
    for (int p = 0; p < max_panels && quiet < 3; ++p) {
        const double mid = (p + 0.5) * h, half = 0.5 * h;
        double d1 = 0.0, d2 = 0.0;
        for (int q = 0; q < 8; ++q)
            for (int s = -1; s <= 1; s += 2) {
                const double phi = mid + s * half * gx[q];
                d1 += gw[q] * half * integrand(phi, 1);
                d2 += gw[q] * half * integrand(phi, 2);
            }
        I1 += d1;
        I2 += d2;
        quiet = (std::fabs(d1) < 1e-16 && std::fabs(d2) < 1e-16) ? quiet + 1 : 0;
    }
    const double pi = 3.14159265358979323846;
    return S0 * (0.5 + I1 / pi) - K * std::exp(-r * T) * (0.5 + I2 / pi);
}

}  // namespace nmch::utils

namespace nmch::utils::cuda {

void checkCUDA(int status, const char *file, int line)
{
    if (status != 0) {
        printf("There is an error in file %s at line %d\n", file, line);
        fprintf(stderr, "nmch_b200: %s: %s\n", nmch_status_string(status), nmch_last_error());
        exit(EXIT_FAILURE);
    }
}

}  // namespace nmch::utils::cuda
