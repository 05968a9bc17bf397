// xorwow_device.cuh -- device pieces of the XORWOW skip-ahead shared by the init kernel (xorwow.cu) and the FE kernels
// that walk a sweep in chunks (fe_kernels.cu).  A skip matrix is 160 rows (one per state bit) of kXorwowRowWords words
// (5 state words padded to 32 bytes: two 16-byte loads per row).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace nmchb {

constexpr int kXorwowRowWords = 8;
constexpr int kXorwowMatWords = 160 * kXorwowRowWords;     // one GF(2) matrix in the table layout

// v <- v * M for a state private to the thread (every lane its own state, and possibly its own matrix).
__device__ __forceinline__ void thread_vecmat(uint32_t v[5], const uint32_t *__restrict__ M)
{
    const uint4 *rows = reinterpret_cast<const uint4 *>(M);
    uint32_t r0 = 0, r1 = 0, r2 = 0, r3 = 0, r4 = 0;
#pragma unroll
    for (int w = 0; w < 5; ++w) {
        const uint32_t word = v[w];
#pragma unroll 8
        for (int j = 0; j < 32; ++j) {
            const uint4 a = __ldg(rows + 2 * (w * 32 + j));
            const uint32_t b4 = __ldg(reinterpret_cast<const uint32_t *>(rows + 2 * (w * 32 + j) + 1));
            const uint32_t mask = 0u - ((word >> j) & 1u);
            r0 ^= a.x & mask; r1 ^= a.y & mask; r2 ^= a.z & mask; r3 ^= a.w & mask; r4 ^= b4 & mask;
        }
    }
    v[0] = r0; v[1] = r1; v[2] = r2; v[3] = r3; v[4] = r4;
}


// Digit walk: v <- v * A^c for tables holding A^(q 4^m), q = 1..3, at [(m * 3 + q - 1) * kXorwowMatWords].
__device__ __forceinline__ void xorwow_apply_digits(uint32_t v[5], const uint32_t *__restrict__ tables, unsigned int c)
{
    for (int m = 0; c != 0u; ++m, c >>= 2) {
        const unsigned int q = c & 3u;
        if (q != 0u) thread_vecmat(v, tables + (size_t)(m * 3 + (int)(q - 1u)) * kXorwowMatWords);
    }
}

}  // namespace nmchb
