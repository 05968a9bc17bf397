/*
 * ref_harness.cu -- a main() of OURS that drives the UNMODIFIED reference classes
 * (/root/reference/include/NMCH/methods/*.hpp, compiled from /root/reference/src by
 * oracle/Makefile into oracle/_ref/nmch_ref_harness).  TEST / BASELINE INFRASTRUCTURE ONLY.
 *
 * Why not the reference's own CLI: src/NMCH/test/nmch.cu hard-wires the Philox instantiation and
 * prints "%f" (6 digits); parity at 1e-5 needs the XORWOW instantiation and full float precision.
 * Everything numerical below is the reference's code: this file only parses flags, calls
 * init/compute through the public API and prints what the getters return.
 *
 *   nmch_ref_harness --method fe|em --rng xorwow|philox|mrg [--kernel k1|k2|k3|k2philox|k1pgm|k1pim] --NTPB .. --NB .. --N ..
 *                    (k1 = NMCH_*_K1_MM: needs a power-of-two NTPB, stores E[X^2]/n^2 in price_squared)
 *                    [--k --theta --sigma --rho --T --S_0 --v_0 --r --seed] [--repeat R]
 *                    [--points FILE]   (lines "k theta sigma": one set_* + compute() per line)
 * One JSON object per compute() on stdout.
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "NMCH/methods/NMCH_EM.hpp"
#include "NMCH/methods/NMCH_FE.hpp"

using namespace nmch::methods;

template <typename Base>
struct Probe : public Base {
    using Base::Base;
    float init_ms() const { return this->Tim_init; }
};

struct Args {
    int NTPB = 512, NB = 512, N = 1000, repeat = 1;
    float T = 1.0f, S_0 = 1.0f, v_0 = 0.1f, r = 0.0f, k = 0.5f, rho = -0.7, theta = 0.1f, sigma = 0.3f;
    unsigned long long seed = 1234;
    std::string method = "fe", rng = "xorwow", kernel = "k3", points;
};

template <typename M>
static int run(const Args &a)
{
    Probe<M> m(a.NTPB, a.NB, a.T, a.S_0, a.v_0, a.r, a.k, a.rho, a.theta, a.sigma, a.N);
    m.init(a.seed);
    std::vector<float> pk, pt, ps;
    if (!a.points.empty()) {
        FILE *f = fopen(a.points.c_str(), "r");
        if (!f) { fprintf(stderr, "cannot open %s\n", a.points.c_str()); return 2; }
        float k, t, s;
        while (fscanf(f, "%f %f %f", &k, &t, &s) == 3) { pk.push_back(k); pt.push_back(t); ps.push_back(s); }
        fclose(f);
    }
    const int n_runs = pk.empty() ? a.repeat : (int)pk.size();
    for (int i = 0; i < n_runs; ++i) {
        float k = a.k, t = a.theta, s = a.sigma;
        if (!pk.empty()) {
            k = pk[i]; t = pt[i]; s = ps[i];
            m.set_k(k); m.set_theta(t); m.set_sigma(s);
        }
        m.compute();
        cudaError_t err = cudaGetLastError();
        printf("{\"impl\": \"reference\", \"method\": \"%s\", \"rng\": \"%s\", \"kernel\": \"%s\", \"NTPB\": %d, \"NB\": %d, "
               "\"N\": %d, \"k\": %.9g, \"theta\": %.9g, \"sigma\": %.9g, \"call\": %d, \"E\": %.9g, \"E2\": %.9g, "
               "\"err\": %.9g, \"exec_ms\": %.6f, \"init_ms\": %.6f, \"cuda\": \"%s\"}\n",
               a.method.c_str(), a.rng.c_str(), a.kernel.c_str(), a.NTPB, a.NB, a.N, k, t, s, i,
               m.get_strike_price(), m.get_price_squared(), m.get_err(), m.get_execution_time(), m.init_ms(),
               cudaGetErrorName(err));
    }
    m.finalize();
    return 0;
}

int main(int argc, char **argv)
{
    Args a;
    for (int i = 1; i < argc; ++i) {
        auto is = [&](const char *f) { return strcmp(argv[i], f) == 0 && i + 1 < argc; };
        if (is("--NTPB")) a.NTPB = atoi(argv[++i]);
        else if (is("--NB")) a.NB = atoi(argv[++i]);
        else if (is("--N")) a.N = atoi(argv[++i]);
        else if (is("--repeat")) a.repeat = atoi(argv[++i]);
        else if (is("--T")) a.T = atof(argv[++i]);
        else if (is("--S_0")) a.S_0 = atof(argv[++i]);
        else if (is("--v_0")) a.v_0 = atof(argv[++i]);
        else if (is("--r")) a.r = atof(argv[++i]);
        else if (is("--k")) a.k = atof(argv[++i]);
        else if (is("--rho")) a.rho = atof(argv[++i]);
        else if (is("--theta")) a.theta = atof(argv[++i]);
        else if (is("--sigma")) a.sigma = atof(argv[++i]);
        else if (is("--seed")) a.seed = strtoull(argv[++i], nullptr, 10);
        else if (is("--method")) a.method = argv[++i];
        else if (is("--rng")) a.rng = argv[++i];
        else if (is("--kernel")) a.kernel = argv[++i];
        else if (is("--points")) a.points = argv[++i];
        else { fprintf(stderr, "unknown flag %s\n", argv[i]); return 2; }
    }
    const bool x = a.rng == "xorwow";
    if (a.rng == "mrg") {
        if (a.method == "fe" && a.kernel == "k1") return run<NMCH_FE_K1_MM<curandStateMRG32k3a_t>>(a);
        if (a.method == "em" && a.kernel == "k1") return run<NMCH_EM_K1_MM<curandStateMRG32k3a_t>>(a);
        if (a.method == "fe" && a.kernel == "k2") return run<NMCH_FE_K2_MM<curandStateMRG32k3a_t>>(a);
        if (a.method == "fe" && a.kernel == "k3") return run<NMCH_FE_K3_MM<curandStateMRG32k3a_t>>(a);
        if (a.method == "em" && a.kernel == "k2") return run<NMCH_EM_K2_MM<curandStateMRG32k3a_t>>(a);
        if (a.method == "em" && a.kernel == "k3") return run<NMCH_EM_K3_MM<curandStateMRG32k3a_t>>(a);
        fprintf(stderr, "unknown method/kernel %s/%s\n", a.method.c_str(), a.kernel.c_str());
        return 2;
    }
    if (!x && a.rng != "philox") { fprintf(stderr, "unknown rng %s\n", a.rng.c_str()); return 2; }
    if (a.method == "fe") {
        if (a.kernel == "k1") return x ? run<NMCH_FE_K1_MM<curandStateXORWOW_t>>(a) : run<NMCH_FE_K1_MM<curandStatePhilox4_32_10_t>>(a);
        if (a.kernel == "k2") return x ? run<NMCH_FE_K2_MM<curandStateXORWOW_t>>(a) : run<NMCH_FE_K2_MM<curandStatePhilox4_32_10_t>>(a);
        if (a.kernel == "k3") return x ? run<NMCH_FE_K3_MM<curandStateXORWOW_t>>(a) : run<NMCH_FE_K3_MM<curandStatePhilox4_32_10_t>>(a);
        if (a.kernel == "k2philox") return run<NMCH_FE_K2_PHILOX_MM>(a);
        // the pageable / pinned host-memory variants of the K1 class (NMCH_FE.hpp:168,180)
        if (a.kernel == "k1pgm") return x ? run<NMCH_FE_K1_PgM<curandStateXORWOW_t>>(a) : run<NMCH_FE_K1_PgM<curandStatePhilox4_32_10_t>>(a);
        if (a.kernel == "k1pim") return x ? run<NMCH_FE_K1_PiM<curandStateXORWOW_t>>(a) : run<NMCH_FE_K1_PiM<curandStatePhilox4_32_10_t>>(a);
    } else if (a.method == "em") {
        if (a.kernel == "k1") return x ? run<NMCH_EM_K1_MM<curandStateXORWOW_t>>(a) : run<NMCH_EM_K1_MM<curandStatePhilox4_32_10_t>>(a);
        if (a.kernel == "k2") return x ? run<NMCH_EM_K2_MM<curandStateXORWOW_t>>(a) : run<NMCH_EM_K2_MM<curandStatePhilox4_32_10_t>>(a);
        if (a.kernel == "k3") return x ? run<NMCH_EM_K3_MM<curandStateXORWOW_t>>(a) : run<NMCH_EM_K3_MM<curandStatePhilox4_32_10_t>>(a);
    }
    fprintf(stderr, "unknown method/kernel %s/%s\n", a.method.c_str(), a.kernel.c_str());
    return 2;
}
