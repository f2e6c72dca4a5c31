// Shared device helpers for libn2v_b200.so (sm_100a). See include/n2v_b200.h for the ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "n2v_b200.h"

namespace n2v {

void set_error(const char *fmt, ...);

#define N2V_CHECK_CUDA(expr)                                                              \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) {                                                          \
            n2v::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),        \
                           __FILE__, __LINE__);                                           \
            return N2V_ECUDA;                                                             \
        }                                                                                 \
    } while (0)

#define N2V_REQUIRE(cond, msg)                                                            \
    do {                                                                                  \
        if (!(cond)) {                                                                    \
            n2v::set_error("%s: %s", __func__, msg);                                      \
            return N2V_EINVAL;                                                            \
        }                                                                                 \
    } while (0)

#define N2V_LAUNCH_CHECK()                                                                \
    do {                                                                                  \
        cudaError_t _e = cudaGetLastError();                                              \
        if (_e != cudaSuccess) {                                                          \
            n2v::set_error("%s: kernel launch failed: %s", __func__, cudaGetErrorString(_e)); \
            return N2V_ECUDA;                                                             \
        }                                                                                 \
    } while (0)

int sm_count();  // cached per device, <0 on error

// ---- Philox4x32-10 (Salmon et al. SC'11); same words as oracle/n2v_oracle.c -----------------
struct Philox4 { uint32_t x, y, z, w; };

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2,
                                                          uint32_t c3, uint32_t k0, uint32_t k1)
{
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
#ifdef __CUDA_ARCH__
        uint32_t hi0 = __umulhi(M0, c0), hi1 = __umulhi(M1, c2);
#else
        uint32_t hi0 = (uint32_t)(((uint64_t)M0 * c0) >> 32), hi1 = (uint32_t)(((uint64_t)M1 * c2) >> 32);
#endif
        uint32_t lo0 = M0 * c0, lo1 = M1 * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += W0; k1 += W1;
    }
    return Philox4{c0, c1, c2, c3};
}

// lower_bound style membership test in a sorted int32 range [lo, hi)
__device__ __forceinline__ bool sorted_contains(const int32_t *__restrict__ col, int64_t lo,
                                                int64_t hi, int32_t x)
{
    while (lo < hi) {
        int64_t mid = lo + ((hi - lo) >> 1);
        int32_t c = __ldg(col + mid);
        if (c == x) return true;
        if (c < x) lo = mid + 1; else hi = mid;
    }
    return false;
}

__device__ __forceinline__ n2v_slot_t make_slot(int32_t k, int32_t J, double q)
{
    // u2 < q  <=>  r2 < ceil(q * 2^32)   (u2 = r2 * 2^-32, exact in float64)
    double t = ceil(q * 4294967296.0);
    n2v_slot_t s;
    if (t >= 4294967296.0) { s.alias = k; s.thr = 0xFFFFFFFFu; }   // always "accept": both branches -> k
    else { s.alias = J; s.thr = (uint32_t)t; }
    return s;
}

}  // namespace n2v
