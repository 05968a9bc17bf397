// fp64_offload.cu -- can the idle FP64 pipe carry part of the Philox multiply?  (pipe_rates2.cu + FP64 classes)
//   WIDEF = a Philox half-round whose HIGH product word comes from the FP64 pipe, exactly:
//     x = lo ^ hi;  nlo = x * M (IMAD, low word);  xd = double(x) (magic-number DADD);
//     u = fma(xd, Ml 2^-32, -(2^20 + nlo 2^-32));  s = fma(xd, Mh 2^-16, u) = hi(x M) - 2^20, both exact (checked on the host
//     with rationals for 2e5 random x and the corner cases);  nhi = low word of s + (2^52 + 2^20)
//   i.e. LOP3 + IMAD + DADD + DFMA + DFMA + DADD in place of LOP3 + IMAD.WIDE.
// pipe_rates2.cu -- round-2 issue-rate microbenchmarks (replaces pipe_rates.cu / fe_mix_bound.cu as the evidence behind
// the FE and EM instruction-mix bounds in DESIGN.md).
//
// What changed against the round-1 tools ("Prove or beat the FE ceiling"):
//   * every chain is truly independent (own registers per instruction class, no class reads another class's output);
//   * the 32x32->64 multiply chain is a Philox half-round, w = (lo(w) ^ hi(w)) * M: ONE LOP3 + ONE IMAD.WIDE.U32, both
//     result words live, nothing to unpack, nothing for ptxas to hoist or strength-reduce (a chain through the low
//     word alone becomes plain IMADs; a chain through the 64-bit addend is split into IMAD.WIDE + IADD3 + IADD3.X).
//     The class "WIDE" therefore costs one ALU-pipe LOP3 per multiply, which the mixes below count as ALU work.  The round-1 form unpacked the product into two separate 32-bit variables, for which
//     ptxas added a MOV / IMAD.MOV per multiply (IMAD.MOV runs on the same pipe) -- that is where "5.2 cycles per
//     IMAD.WIDE" came from; a chain through the low word alone is strength-reduced to plain IMADs by ptxas;
//   * the loop body is unrolled 16x over 8 chains per class, so loop overhead (and the register permutation ptxas adds at the back edge) is < 3 % of the issue slots;
//   * cycles are read on the device (clock64 around the loop, mean and max over blocks), not derived from the nominal
//     clock; the CUDA-event time is printed beside it as a cross-check;
//   * the SASS of every loop body was inspected (cuobjdump) for stray moves: see profiles/r02_pipe_rates2_sass.txt.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rates2 pipe_rates2.cu && ./pipe_rates2 [warps_per_smsp]
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

constexpr int ITER = 256;
constexpr int UNROLL = 16;
constexpr int CH = 8;             // chains per instruction class

enum Op : int { WIDE = 0, LOP, SHF, FFMA3, FFMAI, FMUL, FADD, MUFU_EX2, MUFU_LG2, MUFU_SIN, MUFU_SQRT, MUFU_RSQ, IMADLO, IMADHI, WIDE_RR, I2FP, IADD, FSETSEL, DFMA, DADD, WIDEF, NOPS };

struct Mix {
    int n[NOPS];
    bool interleave;
};

struct Regs {
    unsigned long long w[CH];     // WIDE chains: w = (lo(w) ^ hi(w)) * M
    uint32_t l[CH], l2[CH];       // LOP3 / SHF chains
    uint32_t m[CH];               // IMAD lo / hi chains
    float f[CH], g[CH];           // FP32 chains (f) and MUFU chains (g)
    float c1, c2;
    double d[CH];
    double dc1, dc2;
    uint32_t k1, k2;              // 0x43300000, 0xC1300000: high words of the magic doubles
    uint32_t rm;                  // register multiplier for WIDE_RR
};

template <int OP>
__device__ __forceinline__ void emit(Regs &r, int i)
{
    if constexpr (OP == WIDE)
        asm volatile("{ .reg .u32 lo, hi; mov.b64 {lo, hi}, %0; xor.b32 lo, lo, hi; mul.wide.u32 %0, lo, 0xD2511F53; }" : "+l"(r.w[i]));
    else if constexpr (OP == WIDE_RR)
        asm volatile("{ .reg .u32 lo, hi; mov.b64 {lo, hi}, %0; xor.b32 lo, lo, hi; mul.wide.u32 %0, lo, %1; }" : "+l"(r.w[i]) : "r"(r.rm));
    else if constexpr (OP == LOP)
        asm volatile("lop3.b32 %0, %0, %1, 0x12345678, 0x96;" : "+r"(r.l[i]) : "r"(r.l2[i]));
    else if constexpr (OP == SHF)
        asm volatile("shf.r.wrap.b32 %0, %0, %1, 9;" : "+r"(r.l2[i]) : "r"(r.l[i]));
    else if constexpr (OP == FFMA3)
        asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(r.f[i]) : "f"(r.c1), "f"(r.c2));
    else if constexpr (OP == FFMAI)
        asm volatile("fma.rn.f32 %0, %0, 0f3F7FFFF0, 0f3A000000;" : "+f"(r.f[i]));
    else if constexpr (OP == FMUL)
        asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(r.f[i]) : "f"(r.c1));
    else if constexpr (OP == FADD)
        asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(r.f[i]) : "f"(r.c2));
    else if constexpr (OP == MUFU_EX2)
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(r.g[i]));
    else if constexpr (OP == MUFU_LG2)
        asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(r.g[i]));
    else if constexpr (OP == MUFU_SIN)
        asm volatile("sin.approx.ftz.f32 %0, %0;" : "+f"(r.g[i]));      // FMUL.RZ + MUFU.SIN
    else if constexpr (OP == MUFU_SQRT)
        asm volatile("sqrt.approx.ftz.f32 %0, %0;" : "+f"(r.g[i]));
    else if constexpr (OP == MUFU_RSQ)
        asm volatile("rsqrt.approx.ftz.f32 %0, %0;" : "+f"(r.g[i]));
    else if constexpr (OP == I2FP)       // u32 bits of the chain value -> float (I2FP.F32.U32)
        asm volatile("{ .reg .u32 t; mov.b32 t, %0; cvt.rn.f32.u32 %0, t; }" : "+f"(r.g[i]));
    else if constexpr (OP == IADD)       // LEA: l = (l << 3) + l2 (a plain add chain is folded into one IMAD by ptxas)
        asm volatile("{ .reg .u32 t; shl.b32 t, %0, 3; add.u32 %0, t, %1; }" : "+r"(r.l[i]) : "r"(r.l2[i]));
    else if constexpr (OP == FSETSEL)    // FSETP + FSEL: f = (f < c1) ? c2 : f
        asm volatile("{ .reg .pred p; setp.lt.f32 p, %0, %1; selp.f32 %0, %2, %0, p; }" : "+f"(r.f[i]) : "f"(r.c1), "f"(r.c2));
    else if constexpr (OP == DFMA)
        asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(r.d[i]) : "d"(r.dc1), "d"(r.dc2));
    else if constexpr (OP == DADD)
        asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(r.d[i]) : "d"(r.dc2));
    else if constexpr (OP == WIDEF)
        asm volatile("{ .reg .u32 lo, hi, x, nlo, nhi, junk; .reg .f64 xd, v, u, s, q;\n"
                     "mov.b64 {lo, hi}, %0; xor.b32 x, lo, hi;\n"
                     "mul.lo.u32 nlo, x, 0xD2511F53;\n"
                     "mov.b64 xd, {x, %1}; sub.rn.f64 xd, xd, 0d4330000000000000;\n"
                     "mov.b64 v, {nlo, %2};\n"
                     "fma.rn.f64 u, xd, 0d3EBF530000000000, v;\n"
                     "fma.rn.f64 s, xd, 0d3FEA4A2000000000, u;\n"
                     "add.rn.f64 q, s, 0d4330000000100000;\n"
                     "mov.b64 {nhi, junk}, q; mov.b64 %0, {nlo, nhi}; }" : "+l"(r.w[i]) : "r"(r.k1), "r"(r.k2));
    else if constexpr (OP == IMADLO)
        asm volatile("mad.lo.u32 %0, %0, %0, 0x9E3779B9;" : "+r"(r.m[i]));   // m*m + c: nothing to strength-reduce
    else if constexpr (OP == IMADHI)
        asm volatile("mul.hi.u32 %0, %0, 0xD2511F53;" : "+r"(r.m[i]));
}

// One "step": N0 x OP0, N1 x OP1, N2 x OP2, N3 x OP3, N4 x OP4.  INTERLEAVE: the classes are dealt round-robin in
// proportion (largest remaining deficit first) instead of class after class -- ptxas keeps volatile asm in order, so the
// source order IS the issue order inside a warp.
template <int OP0, int N0, int OP1, int N1, int OP2, int N2, int OP3, int N3, int OP4, int N4, bool INTERLEAVE>
__device__ __forceinline__ void step(Regs &r, int &c0, int &c1, int &c2, int &c3, int &c4)
{
    constexpr int TOTAL = N0 + N1 + N2 + N3 + N4;
    if constexpr (!INTERLEAVE) {
#pragma unroll
        for (int i = 0; i < N0; ++i) emit<OP0>(r, (c0++) % CH);
#pragma unroll
        for (int i = 0; i < N1; ++i) emit<OP1>(r, (c1++) % CH);
#pragma unroll
        for (int i = 0; i < N2; ++i) emit<OP2>(r, (c2++) % CH);
#pragma unroll
        for (int i = 0; i < N3; ++i) emit<OP3>(r, (c3++) % CH);
#pragma unroll
        for (int i = 0; i < N4; ++i) emit<OP4>(r, (c4++) % CH);
    } else {
        int d0 = 0, d1 = 0, d2 = 0, d3 = 0, d4 = 0;       // issued so far, per class
#pragma unroll
        for (int s = 1; s <= TOTAL; ++s) {
            // class with the largest deficit  N_k * s / TOTAL - d_k  (compile-time after unrolling)
            const int e0 = N0 * s - d0 * TOTAL, e1 = N1 * s - d1 * TOTAL, e2 = N2 * s - d2 * TOTAL,
                      e3 = N3 * s - d3 * TOTAL, e4 = N4 * s - d4 * TOTAL;
            int best = 0, be = e0;
            if (e1 > be) { best = 1; be = e1; }
            if (e2 > be) { best = 2; be = e2; }
            if (e3 > be) { best = 3; be = e3; }
            if (e4 > be) { best = 4; be = e4; }
            if (best == 0) { if constexpr (N0 > 0) emit<OP0>(r, (c0++) % CH); ++d0; }
            else if (best == 1) { if constexpr (N1 > 0) emit<OP1>(r, (c1++) % CH); ++d1; }
            else if (best == 2) { if constexpr (N2 > 0) emit<OP2>(r, (c2++) % CH); ++d2; }
            else if (best == 3) { if constexpr (N3 > 0) emit<OP3>(r, (c3++) % CH); ++d3; }
            else { if constexpr (N4 > 0) emit<OP4>(r, (c4++) % CH); ++d4; }
        }
    }
}

template <int OP0, int N0, int OP1, int N1, int OP2, int N2, int OP3, int N3, int OP4, int N4, bool INTERLEAVE>
__global__ void __launch_bounds__(1024) mix_kernel(long long *cycles, uint32_t *sink, uint32_t seed)
{
    Regs r;
#pragma unroll
    for (int i = 0; i < CH; ++i) {
        r.w[i] = ((unsigned long long)(seed + i) << 32) | (threadIdx.x * 7u + i + 1u);
        r.l[i] = seed + threadIdx.x * 3 + i;
        r.l2[i] = seed ^ (i * 0x9E3779B9u);
        r.m[i] = seed + threadIdx.x + i * 5 + 1;
        r.f[i] = 1.0f + (float)(threadIdx.x + i) * 1e-3f;
        r.g[i] = 1.25f + (float)i * 0.125f;
    }
    r.c1 = 0.99999f + (float)seed * 1e-9f;
    r.c2 = 1e-3f * (float)seed;
    r.rm = 0xD2511F53u + seed * 2u;
#pragma unroll
    for (int i = 0; i < CH; ++i) r.d[i] = 1.0 + (double)(threadIdx.x + i) * 1e-3;
    r.dc1 = 0.99999 + (double)seed * 1e-9;
    r.dc2 = 1e-3 * (double)seed;
    r.k1 = 0x43300000u + (seed >> 8);
    r.k2 = 0xC1300000u + (seed >> 8);
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
        int c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0;
#pragma unroll
        for (int u = 0; u < UNROLL; ++u)
            step<OP0, N0, OP1, N1, OP2, N2, OP3, N3, OP4, N4, INTERLEAVE>(r, c0, c1, c2, c3, c4);
    }
    const long long t1 = clock64();
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i)
        x ^= (uint32_t)r.w[i] ^ (uint32_t)(r.w[i] >> 32) ^ r.l[i] ^ r.l2[i] ^ r.m[i] ^ __float_as_uint(r.f[i]) ^ __float_as_uint(r.g[i]) ^ (uint32_t)__double_as_longlong(r.d[i]);
    if (x == 0x12345u) sink[0] = x;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

static int g_warps_per_smsp = 12;
static int g_sms = 148;
static double g_ghz = 1.965;
static long long *g_cycles;
static uint32_t *g_sink;

template <int OP0, int N0, int OP1, int N1, int OP2, int N2, int OP3, int N3, int OP4, int N4, bool INTERLEAVE>
void run(const char *name)
{
    // one block per SM, warps_per_smsp * 4 warps: every block is resident from the start, one wave
    const int threads = g_warps_per_smsp * 4 * 32;
    auto kern = mix_kernel<OP0, N0, OP1, N1, OP2, N2, OP3, N3, OP4, N4, INTERLEAVE>;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    kern<<<g_sms, threads>>>(g_cycles, g_sink, 1);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    kern<<<g_sms, threads>>>(g_cycles, g_sink, 2);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    std::vector<long long> h(g_sms);
    cudaMemcpy(h.data(), g_cycles, g_sms * sizeof(long long), cudaMemcpyDeviceToHost);
    double mean = 0, mx = 0;
    for (long long c : h) { mean += (double)c; if ((double)c > mx) mx = (double)c; }
    mean /= g_sms;
    const int total = N0 + N1 + N2 + N3 + N4;
    const double steps_per_smsp = (double)ITER * UNROLL * g_warps_per_smsp;   // warp-steps issued by one SMSP
    const double cyc_step = mean / steps_per_smsp;
    cudaError_t err = cudaGetLastError();
    printf("%-64s %2d instr  %7.2f cyc/warp-step/SMSP  (%5.3f cyc/instr, %5.1f thread-instr/clk/SM; max-block %7.2f; event %.3f ms = %.2f cyc at %.3f GHz)%s\n",
           name, total, cyc_step, cyc_step / total, 128.0 * total / cyc_step, mx / steps_per_smsp, ms,
           ms * 1e-3 * g_ghz * 1e9 / steps_per_smsp, g_ghz, err == cudaSuccess ? "" : cudaGetErrorString(err));
}

#define RUN1(name, OP, N) run<OP, N, LOP, 0, LOP, 0, LOP, 0, LOP, 0, false>(name)
#define RUN2(name, A, NA, B, NB, IL) run<A, NA, B, NB, LOP, 0, LOP, 0, LOP, 0, IL>(name)
#define RUN3(name, A, NA, B, NB, C, NC, IL) run<A, NA, B, NB, C, NC, LOP, 0, LOP, 0, IL>(name)

int main(int argc, char **argv)
{
    if (argc > 1) g_warps_per_smsp = atoi(argv[1]);
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    g_sms = p.multiProcessorCount;
    g_ghz = khz * 1e-6;
    cudaMalloc(&g_cycles, g_sms * sizeof(long long));
    cudaMalloc(&g_sink, 4);
    printf("%s, %d SMs, %.3f GHz nominal, %d warps per SMSP, %d chains per class, unroll %d\n", p.name, g_sms, g_ghz,
           g_warps_per_smsp, CH, UNROLL);
    // exactness on the device: WIDEF against mul.wide on 2^20 inputs per thread is checked by check_kernel below
    printf("-- single classes (8 per step)\n");
    RUN1("FFMA 3-reg", FFMA3, 8);
    RUN1("LOP3", LOP, 8);
    RUN1("IMAD.WIDE.U32 + LOP3 (Philox half-round)", WIDE, 8);
    RUN1("IMAD (mad.lo)", IMADLO, 8);
    RUN1("DFMA", DFMA, 8);
    RUN1("DADD", DADD, 8);
    RUN1("half-round with the high word from the FP64 pipe (LOP3 + IMAD + 2 DADD + 2 DFMA)", WIDEF, 8);
    printf("-- pairs\n");
    RUN2("8 DFMA + 8 FFMA3, interleaved", DFMA, 8, FFMA3, 8, true);
    RUN2("8 DFMA + 8 LOP3, interleaved", DFMA, 8, LOP, 8, true);
    RUN2("4 DFMA + 4 WIDE, interleaved", DFMA, 4, WIDE, 4, true);
    RUN2("8 DFMA + 4 WIDE, interleaved", DFMA, 8, WIDE, 4, true);
    RUN2("16 DFMA + 4 WIDE, interleaved", DFMA, 16, WIDE, 4, true);
    RUN2("4 DFMA + 2 MUFU.EX2, interleaved", DFMA, 4, MUFU_EX2, 2, true);
    RUN2("6 WIDE + 2 WIDEF, interleaved", WIDE, 6, WIDEF, 2, true);
    RUN2("4 WIDE + 4 WIDEF, interleaved", WIDE, 4, WIDEF, 4, true);
    printf("-- FE native mix per path-step: 8 half-rounds, 3 further ALU, 12 FP32, 4 MUFU; k of the 8 half-rounds through the FP64 pipe\n");
    run<WIDE, 8, LOP, 3, FFMA3, 12, MUFU_EX2, 4, WIDEF, 0, true>("FE mix, 0 of 8 through FP64");
    run<WIDE, 7, LOP, 3, FFMA3, 12, MUFU_EX2, 4, WIDEF, 1, true>("FE mix, 1 of 8 through FP64");
    run<WIDE, 6, LOP, 3, FFMA3, 12, MUFU_EX2, 4, WIDEF, 2, true>("FE mix, 2 of 8 through FP64");
    run<WIDE, 5, LOP, 3, FFMA3, 12, MUFU_EX2, 4, WIDEF, 3, true>("FE mix, 3 of 8 through FP64");
    run<WIDE, 4, LOP, 3, FFMA3, 12, MUFU_EX2, 4, WIDEF, 4, true>("FE mix, 4 of 8 through FP64");
    run<WIDE, 6, LOP, 3, FFMA3, 12, MUFU_EX2, 4, WIDEF, 2, false>("FE mix, 2 of 8 through FP64, grouped");
    printf("-- the same with the FP64 work as plain independent DFMAs (upper bound of what overlap could give)\n");
    run<WIDE, 6, LOP, 3, FFMA3, 12, MUFU_EX2, 4, DFMA, 8, true>("FE mix with 6 half-rounds + 8 DFMA");
    run<WIDE, 6, LOP, 3, FFMA3, 12, MUFU_EX2, 4, DFMA, 0, true>("FE mix with 6 half-rounds");
    return 0;
}
