// Random-access HBM roofline denominators: the two access shapes of the hot path,
// stripped of everything else.
//   mode 0: independent random 32-byte sector reads (the walk's col / slot / row_ptr gathers)
//   mode 1: random 512-byte row read-modify-write, one row per warp instruction (SGNS rows)
//   mode 2: random 512-byte row read            } the SGNS row traffic split into its halves; buf may be
//   mode 3: random 512-byte row red.add.v4.f32  } a peer GPU's memory (NVLink), which is what these
//   mode 4: random 512-byte row red.add.f32 x4  } three are for
#include "n2v_common.cuh"

namespace n2v {

__device__ __forceinline__ uint64_t mix64(uint64_t z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void __launch_bounds__(256)
gather_sector_kernel(const uint4 *__restrict__ buf, uint64_t n_sectors, int64_t n_access, uint64_t seed,
                     unsigned long long *__restrict__ sink)
{
    const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    uint32_t acc = 0;
    for (int64_t i = tid; i < n_access; i += 4 * stride) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {       // 4 independent sectors in flight per thread
            const int64_t j = i + u * stride;
            const uint64_t sct = mix64(seed + (uint64_t)j) % n_sectors;
            v[u] = j < n_access ? __ldg(buf + sct * 2) : make_uint4(0, 0, 0, 0);   // 16 B of a 32 B sector
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) acc += v[u].x ^ v[u].w;
    }
    if (acc == 0x12345678u) atomicAdd(sink, 1ull);
}

__global__ void __launch_bounds__(256)
row_rmw_kernel(float4 *__restrict__ buf, uint64_t n_rows, int64_t n_access, uint64_t seed)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp; i < n_access; i += 4 * n_warps) {
        float4 v[4]; uint64_t row[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t j = i + u * n_warps;
            row[u] = mix64(seed + (uint64_t)j) % n_rows;
            if (j < n_access) v[u] = buf[row[u] * 32 + lane];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t j = i + u * n_warps;
            if (j < n_access) { v[u].x += 1.0f; v[u].w -= 1.0f; buf[row[u] * 32 + lane] = v[u]; }
        }
    }
}

template <int MODE>
__global__ void __launch_bounds__(256)
row_half_kernel(float4 *__restrict__ buf, uint64_t n_rows, int64_t n_access, uint64_t seed,
                unsigned long long *__restrict__ sink)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    float acc = 0.f;
    for (int64_t i = warp; i < n_access; i += 4 * n_warps) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t j = i + u * n_warps;
            if (j >= n_access) continue;
            float4 *a = buf + (mix64(seed + (uint64_t)j) % n_rows) * 32 + lane;
            if (MODE == 2) {
                v[u] = __ldcg(a);
                acc += v[u].x + v[u].w;
            } else if (MODE == 3) {
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(a), "f"(1.f), "f"(0.f), "f"(0.f), "f"(-1.f) : "memory");
            } else {
                float *f = (float *)a;
                atomicAdd(f, 1.f); atomicAdd(f + 1, 0.f); atomicAdd(f + 2, 0.f); atomicAdd(f + 3, -1.f);
            }
        }
    }
    if (MODE == 2 && acc == 1.2345e30f) atomicAdd(sink, 1ull);
}

}  // namespace n2v

using namespace n2v;

extern "C" int n2v_random_gather_bench(void *buf, size_t n_bytes, int64_t n_access, int mode,
                                       uint64_t seed, unsigned long long *sink, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    N2V_REQUIRE(buf && n_bytes >= 4096 && n_access > 0, "bad argument");
    int sms = sm_count();
    if (sms <= 0) { set_error("no CUDA device"); return N2V_ECUDA; }
    const int blocks = sms * 8;
    if (mode == 0) {
        N2V_REQUIRE(sink, "sink is NULL");
        gather_sector_kernel<<<blocks, 256, 0, stream>>>((const uint4 *)buf, n_bytes / 32, n_access, seed, sink);
    } else if (mode == 1) {
        row_rmw_kernel<<<blocks, 256, 0, stream>>>((float4 *)buf, n_bytes / 512, n_access, seed);
    } else if (mode == 2) {
        N2V_REQUIRE(sink, "sink is NULL");
        row_half_kernel<2><<<blocks, 256, 0, stream>>>((float4 *)buf, n_bytes / 512, n_access, seed, sink);
    } else if (mode == 3) {
        row_half_kernel<3><<<blocks, 256, 0, stream>>>((float4 *)buf, n_bytes / 512, n_access, seed, sink);
    } else if (mode == 4) {
        row_half_kernel<4><<<blocks, 256, 0, stream>>>((float4 *)buf, n_bytes / 512, n_access, seed, sink);
    } else {
        set_error("n2v_random_gather_bench: unknown mode %d", mode);
        return N2V_EINVAL;
    }
    N2V_LAUNCH_CHECK();
    return N2V_OK;
}
