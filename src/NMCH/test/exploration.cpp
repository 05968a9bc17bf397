// exploration.cpp -- the `exploration` tool: the (k, theta, sigma) sweep of the reference
// (src/NMCH/test/exploration.cu:21-120) with the same grid, skip filter, CSV header and row format -- but each
// method's whole sweep is ONE batched launch instead of ~200 launch+sync round trips.
//
// No arguments = the reference's run (NTPB 512, NB 10, N 1000, XORWOW tag, seed 1234, 6x6x6 float-accumulated
// grid, filter 20*k*theta < sigma^2).  Additive flags:
//   --points P          P points per axis on an evenly spaced grid computed in double (P = 20 is BASELINE configs[3])
//   --log2-paths L      2^L paths per grid point (default: NTPB*NB = 5120)
//   --method fe|em|both (default both)   --rng xorwow|xorwow-fast|philox|philox-compat|philox-dense (default xorwow, as the reference)
//   --no-filter         keep the points the reference skips      --gpus N      --N steps   --seed s
//   --bias              extra CSV column: estimate minus the semi-analytic Heston price (input of heatmap.py)
//   --sequential        per-point launches like the reference (per-point execution_time is then measured, not averaged)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include "NMCH/methods/NMCH_EM.hpp"
#include "NMCH/methods/NMCH_FE.hpp"

using namespace nmch::methods;

namespace {

struct Options {
    int NTPB = 512, NB = 10, N = 1000, points = 0, gpus = 1, log2_paths = -1;
    unsigned long long seed = 1234;
    std::string method = "both", rng = "xorwow";
    bool filter = true, bias = false, sequential = false;
};

struct Grid {
    std::vector<float> k, theta, sigma;
};

Grid make_grid(const Options &o)
{
    const float k_min = 0.1f, k_max = 10.0f, theta_min = 0.01f, theta_max = 0.5f, sigma_min = 0.1f, sigma_max = 1.0f;
    Grid g;
    auto keep = [&](float k, float theta, float sigma) {
        if (o.filter && (20 * k * theta < sigma * sigma)) return;      // exploration.cu:76
        g.k.push_back(k); g.theta.push_back(theta); g.sigma.push_back(sigma);
    };
    if (o.points <= 0) {
        // the reference's float-accumulated loops (exploration.cu:46-52, 71-73): sigma outer, theta, k inner
        const float sigma_step = (sigma_max - sigma_min) / 5, theta_step = (theta_max - theta_min) / 5, k_step = (k_max - k_min) / 5;
        for (float sigma = sigma_min; sigma <= sigma_max; sigma += sigma_step)
            for (float theta = theta_min; theta <= theta_max; theta += theta_step)
                for (float k = k_min; k <= k_max; k += k_step) keep(k, theta, sigma);
    } else {
        const int P = o.points;
        auto at = [&](double lo, double hi, int i) { return (float)(P > 1 ? lo + i * (hi - lo) / (P - 1) : lo); };
        for (int l = 0; l < P; ++l)
            for (int j = 0; j < P; ++j)
                for (int i = 0; i < P; ++i) keep(at(0.1, 10.0, i), at(0.01, 0.5, j), at(0.1, 1.0, l));
    }
    return g;
}

template <typename M>
void sweep(const char *name, const Options &o, const Grid &g)
{
    const float T = 1.0f, S_0 = 1.0f, v_0 = 0.1f, r = 0.0f, rho = -0.7;
    M m(o.NTPB, o.NB, T, S_0, v_0, r, 0.5f, rho, 0.1f, 0.3f, o.N);
    m.set_gpus(o.gpus);
    m.set_philox_compat(o.rng == "philox-compat");
    m.set_philox_dense(o.rng == "philox-dense");      // FE-only opt-in streams: the EM leg of the sweep ignores them
    m.set_xorwow_fast(o.rng == "xorwow-fast");
    m.init(o.seed);
    m.compute();                      // the reference's warm-up compute (exploration.cu:65-67); it advances the streams
    const int n = (int)g.k.size();
    std::vector<float> e1(n), e2(n), err(n), ms(n);
    if (o.sequential) {
        for (int i = 0; i < n; ++i) {
            m.set_theta(g.theta[i]); m.set_sigma(g.sigma[i]); m.set_k(g.k[i]);
            m.compute();
            e1[i] = m.get_strike_price(); e2[i] = m.get_price_squared(); err[i] = m.get_err(); ms[i] = m.get_execution_time();
        }
    } else {
        const float total = m.compute_grid(n, g.k.data(), g.theta.data(), g.sigma.data(), e1.data(), e2.data(), err.data());
        for (int i = 0; i < n; ++i) ms[i] = total / n;     // one launch: the per-point time is the launch time / points
    }
    std::vector<double> ref(o.bias ? n : 0);
    if (o.bias) {                                     // semi-analytic prices on all host threads
        const unsigned nt = std::max(1u, std::thread::hardware_concurrency());
        std::vector<std::thread> pool;
        for (unsigned t = 0; t < nt; ++t)
            pool.emplace_back([&, t] {
                for (int i = (int)t; i < n; i += (int)nt)
                    ref[i] = nmch::utils::heston_call(S_0, S_0, v_0, r, g.k[i], g.theta[i], g.sigma[i], rho, T);
            });
        for (auto &th : pool) th.join();
    }
    for (int i = 0; i < n; ++i) {
        if (o.bias) {
            printf("%s, %f, %f, %f, %f, %f, %f\n", name, g.k[i], g.theta[i], g.sigma[i], ms[i], err[i], e1[i] - (float)ref[i]);
        } else {
            printf("%s, %f, %f, %f, %f, %f\n", name, g.k[i], g.theta[i], g.sigma[i], ms[i], err[i]);
        }
    }
    m.finalize();
}

}  // namespace

int main(int argc, char **argv)
{
    Options o;
    for (int i = 1; i < argc; ++i) {
        auto has = [&](const char *f) { return strcmp(argv[i], f) == 0 && i + 1 < argc; };
        if (has("--points")) o.points = atoi(argv[++i]);
        else if (has("--log2-paths")) o.log2_paths = atoi(argv[++i]);
        else if (has("--method")) o.method = argv[++i];
        else if (has("--rng")) o.rng = argv[++i];
        else if (has("--gpus")) o.gpus = atoi(argv[++i]);
        else if (has("--N")) o.N = atoi(argv[++i]);
        else if (has("--seed")) o.seed = strtoull(argv[++i], nullptr, 10);
        else if (strcmp(argv[i], "--no-filter") == 0) o.filter = false;
        else if (strcmp(argv[i], "--bias") == 0) o.bias = true;
        else if (strcmp(argv[i], "--sequential") == 0) o.sequential = true;
        else if (strcmp(argv[i], "--help") == 0) {
            printf("Usage: %s [--points P] [--log2-paths L] [--method fe|em|both] [--rng xorwow|xorwow-fast|philox|philox-compat|philox-dense]\n"
                   "          [--no-filter] [--gpus N] [--N steps] [--seed s] [--bias] [--sequential]\n", argv[0]);
            return 0;
        } else {
            printf("Unknown option: %s\n", argv[i]);
            return 1;
        }
    }
    if (o.log2_paths >= 9) { o.NTPB = 512; o.NB = 1 << (o.log2_paths - 9); }
    const Grid g = make_grid(o);
    const bool x = o.rng == "xorwow" || o.rng == "xorwow-fast";
    if (!x && o.rng != "philox" && o.rng != "philox-compat" && o.rng != "philox-dense") {
        printf("Unknown rng: %s\n", o.rng.c_str());
        return 1;
    }
    printf(o.bias ? "method, k, theta, sigma, execution_time, err, bias\n" : "method, k, theta, sigma, execution_time, err\n");
    if (o.method == "fe" || o.method == "both") {
        if (x) sweep<NMCH_FE_K3_MM<curandStateXORWOW_t>>("fe", o, g);
        else sweep<NMCH_FE_K3_MM<curandStatePhilox4_32_10_t>>("fe", o, g);
    }
    if (o.method == "em" || o.method == "both") {
        if (x) sweep<NMCH_EM_K3_MM<curandStateXORWOW_t>>("em", o, g);
        else sweep<NMCH_EM_K3_MM<curandStatePhilox4_32_10_t>>("em", o, g);
    }
    if (o.method != "fe" && o.method != "em" && o.method != "both") { printf("Unknown method: %s\n", o.method.c_str()); return 1; }
    return 0;
}
