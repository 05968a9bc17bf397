// xorwow.cu -- builds per-path XORWOW generator state compatible with cuRAND's
// curand_init(seed, subsequence = global path index, offset = 0) (reference: src/NMCH/random/random.cu:6-10).
//
// cuRAND ships its skip-ahead matrices as tables; this engine derives them from the generator's
// definition: the five-word xorshift part of XORWOW is linear over GF(2), so advancing by 2^67 draws is
// the 160x160 bit matrix T^(2^67) (67 squarings of the one-step matrix T), and digit m of the
// subsequence in base 4 applies (T^(2^67))^(4^m) q times.  We store the q = 1, 2, 3 powers per digit so
// a digit costs at most one vector-matrix product instead of up to three.
#include <cstring>
#include <mutex>
#include <vector>

#include "kernels.cuh"
#include "xorwow_device.cuh"

namespace nmchb {

namespace {

constexpr int kDigits = 32;            // 64-bit subsequence, 2 bits per digit
constexpr int kRowWords = kXorwowRowWords;   // 5 state words padded to 32 bytes: two 16-byte loads per row

struct Gf2Mat {
    uint32_t row[160][5];              // row b = image of unit vector e_b
};

void host_step(uint32_t v[5])
{
    const uint32_t t = v[0] ^ (v[0] >> 2);
    v[0] = v[1]; v[1] = v[2]; v[2] = v[3]; v[3] = v[4];
    v[4] = (v[4] ^ (v[4] << 4)) ^ (t ^ (t << 1));
}

void host_vecmat(const uint32_t v[5], const Gf2Mat &M, uint32_t out[5])
{
    // walk the set bits only (no data-dependent branch per bit: the whole table derivation takes 4 ms instead of 40)
    uint32_t r0 = 0, r1 = 0, r2 = 0, r3 = 0, r4 = 0;
    for (int w = 0; w < 5; ++w) {
        uint32_t bits = v[w];
        while (bits) {
            const uint32_t *row = M.row[w * 32 + __builtin_ctz(bits)];
            bits &= bits - 1u;
            r0 ^= row[0]; r1 ^= row[1]; r2 ^= row[2]; r3 ^= row[3]; r4 ^= row[4];
        }
    }
    out[0] = r0; out[1] = r1; out[2] = r2; out[3] = r3; out[4] = r4;
}

void host_matmul(const Gf2Mat &A, const Gf2Mat &B, Gf2Mat &out)   // out = A*B (apply A, then B)
{
    Gf2Mat tmp;
    for (int b = 0; b < 160; ++b) host_vecmat(A.row[b], B, tmp.row[b]);
    out = tmp;
}

}  // namespace

struct XorwowSkipTables {
    uint32_t *d_tables = nullptr;      // [kDigits][3][160][kRowWords]
};

static const Gf2Mat &one_step_matrix()
{
    static Gf2Mat T;
    static std::once_flag once;
    std::call_once(once, [] {
        for (int b = 0; b < 160; ++b) {
            uint32_t e[5] = {0, 0, 0, 0, 0};
            e[b >> 5] = 1u << (b & 31);
            host_step(e);
            std::memcpy(T.row[b], e, sizeof e);
        }
    });
    return T;
}

// Offset skip-ahead (what cuRAND's curand_init does for its `offset` argument with its own tables,
// curand_kernel.h:703-719): A = T^draws by square-and-multiply
// from the one-step matrix, then A^(q 4^m) for q = 1..3, m < n_digits, in the device table layout.  The FE kernels use
// it to start chunk c of a sweep c * draws positions into each path's stream (fe_kernels.cu).
std::vector<uint32_t> xorwow_offset_tables_host(unsigned long long draws, int n_digits)
{
    Gf2Mat acc, sq = one_step_matrix();
    bool have = false;
    for (unsigned long long e = draws; e != 0ull; e >>= 1) {
        if (e & 1ull) {
            if (have) host_matmul(acc, sq, acc);
            else { acc = sq; have = true; }
        }
        if (e >> 1) host_matmul(sq, sq, sq);
    }
    if (!have)                                                   // draws == 0: identity
        for (int b = 0; b < 160; ++b) {
            std::memset(acc.row[b], 0, sizeof acc.row[b]);
            acc.row[b][b >> 5] = 1u << (b & 31);
        }
    std::vector<uint32_t> host((size_t)n_digits * 3 * 160 * kRowWords, 0u);
    Gf2Mat cur = acc;
    for (int m = 0; m < n_digits; ++m) {
        Gf2Mat p2, p3;
        host_matmul(cur, cur, p2);
        host_matmul(p2, cur, p3);
        const Gf2Mat *pw[3] = {&cur, &p2, &p3};
        for (int q = 0; q < 3; ++q)
            for (int b = 0; b < 160; ++b)
                std::memcpy(&host[(((size_t)m * 3 + q) * 160 + b) * kRowWords], pw[q]->row[b], 5 * sizeof(uint32_t));
        if (m + 1 < n_digits) host_matmul(p2, p2, cur);              // next digit: fourth power
    }
    return host;
}

static const std::vector<uint32_t> &host_tables()
{
    static std::vector<uint32_t> host;
    static std::once_flag once;
    std::call_once(once, [] {
    Gf2Mat cur;
    for (int b = 0; b < 160; ++b) {
        uint32_t e[5] = {0, 0, 0, 0, 0};
        e[b >> 5] = 1u << (b & 31);
        host_step(e);
        std::memcpy(cur.row[b], e, sizeof e);
    }
    for (int s = 0; s < 67; ++s) host_matmul(cur, cur, cur);           // T^(2^67)
    host.assign((size_t)kDigits * 3 * 160 * kRowWords, 0u);
    for (int m = 0; m < kDigits; ++m) {
        Gf2Mat p2, p3;
        host_matmul(cur, cur, p2);
        host_matmul(p2, cur, p3);
        const Gf2Mat *pw[3] = {&cur, &p2, &p3};
        for (int q = 0; q < 3; ++q)
            for (int b = 0; b < 160; ++b)
                std::memcpy(&host[(((size_t)m * 3 + q) * 160 + b) * kRowWords], pw[q]->row[b], 5 * sizeof(uint32_t));
        host_matmul(p2, p2, cur);                                      // next digit: fourth power
    }
    });
    return host;
}

cudaError_t xorwow_tables_create(XorwowSkipTables **out)
{
    const std::vector<uint32_t> &host = host_tables();
    auto *t = new XorwowSkipTables();
    cudaError_t err = cudaMalloc(&t->d_tables, host.size() * sizeof(uint32_t));
    if (err == cudaSuccess)
        err = cudaMemcpy(t->d_tables, host.data(), host.size() * sizeof(uint32_t), cudaMemcpyHostToDevice);
    if (err != cudaSuccess) {
        xorwow_tables_destroy(t);
        return err;
    }
    *out = t;
    return cudaSuccess;
}

void xorwow_tables_destroy(XorwowSkipTables *t)
{
    if (!t) return;
    if (t->d_tables) cudaFree(t->d_tables);
    delete t;
}

// v <- v * M for ONE state held replicated by all 32 lanes of a warp: lane l owns bit l of each of the five
// words (5 rows, two 16-byte loads each), then the partial XORs are folded with shuffles.
__device__ __forceinline__ void warp_vecmat(uint32_t v[5], const uint32_t *__restrict__ M, int lane)
{
    uint32_t r[5] = {0u, 0u, 0u, 0u, 0u};
#pragma unroll
    for (int w = 0; w < 5; ++w) {
        const uint4 *row = reinterpret_cast<const uint4 *>(M + (size_t)(w * 32 + lane) * kRowWords);
        const uint4 a = __ldg(row);
        const uint32_t b4 = __ldg(reinterpret_cast<const uint32_t *>(row + 1));
        const uint32_t mask = 0u - ((v[w] >> lane) & 1u);
        r[0] ^= a.x & mask; r[1] ^= a.y & mask; r[2] ^= a.z & mask; r[3] ^= a.w & mask; r[4] ^= b4 & mask;
    }
#pragma unroll
    for (int k = 0; k < 5; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r[k] ^= __shfl_xor_sync(0xffffffffu, r[k], o);
        v[k] = r[k];
    }
}

// Block of 256 consecutive paths.  When the block's first path is a multiple of 256 all its paths share the
// digits above bit 8 of the subsequence: warp 0 applies those once for the whole block (cooperative products),
// and every thread then applies only its own four low digits -- about a third of the work of a full walk.
__global__ void __launch_bounds__(256)
xorwow_init_kernel(const uint32_t *__restrict__ tables, unsigned long long seed, unsigned long long first_path,
                   unsigned long long n_local, XorwowState xs)
{
    __shared__ uint32_t s_base[5];
    const unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    // seed scramble of XORWOW as cuRAND defines it (curand_kernel.h:807-818)
    const uint32_t s0 = (uint32_t)seed ^ 0xaad26b49u;
    const uint32_t s1 = (uint32_t)(seed >> 32) ^ 0xf7dcefddu;
    const uint32_t t0 = 1099087573u * s0;
    const uint32_t t1 = 2591861531u * s1;
    uint32_t v[5] = {123456789u + t0, 362436069u ^ t0, 521288629u + t1, 88675123u ^ t1, 5783321u + t0};
    const uint32_t d = 6615241u + t1 + t0;

    const unsigned long long block_first = first_path + (unsigned long long)blockIdx.x * blockDim.x;
    const bool aligned = (block_first & 255ull) == 0ull;          // uniform over the block
    unsigned long long sub = first_path + idx;
    int m = 0;
    if (aligned) {
        if (threadIdx.x < 32) {
            unsigned long long hi = block_first >> 8;
            for (int mm = 4; hi != 0ull; ++mm, hi >>= 2) {
                const unsigned q = (unsigned)(hi & 3ull);
                NMCHB_ASSERT(mm < kDigits);
                if (q != 0u) warp_vecmat(v, tables + ((size_t)mm * 3 + (q - 1)) * 160 * kRowWords, threadIdx.x);
            }
            if (threadIdx.x < 5) s_base[threadIdx.x] = v[threadIdx.x];
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 5; ++k) v[k] = s_base[k];
        sub &= 255ull;                                            // the digits still to apply
    }
    if (idx >= n_local) return;
    for (; sub != 0ull; ++m, sub >>= 2) {
        const unsigned q = (unsigned)(sub & 3ull);
        NMCHB_ASSERT(m < kDigits);
        if (q != 0u) thread_vecmat(v, tables + ((size_t)m * 3 + (q - 1)) * 160 * kRowWords);
    }
    xs.d[idx] = d;
    xs.v0[idx] = v[0]; xs.v1[idx] = v[1]; xs.v2[idx] = v[2]; xs.v3[idx] = v[3]; xs.v4[idx] = v[4];
    if (xs.bm_flag) {
        xs.bm_flag[idx] = 0;
        xs.bm_extra[idx] = 0.0f;
        xs.bm_flag_d[idx] = 0;
        xs.bm_extra_d[idx] = 0.0;
    }
}

cudaError_t launch_xorwow_init(const XorwowSkipTables *t, unsigned long long seed, unsigned long long first_path,
                               unsigned long long n_local, XorwowState xs, cudaStream_t stream)
{
    if (n_local == 0) return cudaSuccess;
    const unsigned long long blocks = (n_local + 255ull) / 256ull;
    xorwow_init_kernel<<<(unsigned)blocks, 256, 0, stream>>>(t->d_tables, seed, first_path, n_local, xs);
    return cudaGetLastError();
}

}  // namespace nmchb
