// Random-access HBM roofline denominators: the two access shapes of the hot path,
// stripped of everything else.
//   mode 0: independent random 32-byte sector reads (the walk's col / slot / row_ptr gathers)
//   mode 1: random 512-byte row read-modify-write, one row per warp instruction (SGNS rows)
//   mode 2: random 512-byte row read            } the SGNS row traffic split into its halves; buf may be
//   mode 3: random 512-byte row red.add.v4.f32  } a peer GPU's memory (NVLink), which is what these
//   mode 4: random 512-byte row red.add.f32 x4  } three are for
//   mode 5: random 512-byte row read by cp.async.bulk (global -> shared, mbarrier complete_tx; UBLKCP)
//   mode 6: random 512-byte row reduction by cp.reduce.async.bulk.add.f32 (shared -> global; UBLKRED)
//   mode 7: mode 6 with the source row rewritten (st.shared + fence.proxy.async) before every reduction
//   mode 8: mode 0 with 16 independent sector reads in flight per thread (validated against ncu's
//           dram__sectors_read: profiles/r02_*)
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

// ---- bulk-async forms of the row traffic (TMA engine instead of the LSU) ---------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

constexpr int BULK_DEPTH = 8;      // rows in flight per warp
constexpr int BULK_WARPS = 4;      // 4 warps x 8 rows x 512 B = 16 KB per block

__global__ void __launch_bounds__(BULK_WARPS * 32)
row_bulk_read_kernel(const float4 *__restrict__ buf, uint64_t n_rows, int64_t n_access, uint64_t seed,
                     unsigned long long *__restrict__ sink)
{
    __shared__ __align__(128) float4 s_row[BULK_WARPS][BULK_DEPTH][32];
    __shared__ __align__(8) uint64_t s_bar[BULK_WARPS][BULK_DEPTH];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    if (lane < BULK_DEPTH)
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar[wib][lane])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    auto issue = [&](int64_t k, int slot) {         // lane 0: one 512-byte bulk copy
        const uint64_t row = mix64(seed + (uint64_t)k) % n_rows;
        const uint32_t bar = smem_u32(&s_bar[wib][slot]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 512;" ::"r"(bar) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 512, [%2];"
                     ::"r"(smem_u32(&s_row[wib][slot][0])), "l"(buf + row * 32), "r"(bar) : "memory");
    };
    float acc = 0.f;
    int64_t k = warp;
    int64_t issued = 0;
    for (int d = 0; d < BULK_DEPTH; ++d, ++issued) {
        const int64_t kk = warp + issued * n_warps;
        if (kk < n_access && lane == 0) issue(kk, d);
    }
    int64_t done = 0;
    for (; k < n_access; k += n_warps, ++done) {
        const int slot = (int)(done % BULK_DEPTH);
        const uint32_t parity = (uint32_t)((done / BULK_DEPTH) & 1);
        const uint32_t bar = smem_u32(&s_bar[wib][slot]);
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        const float4 v = s_row[wib][slot][lane];
        acc += v.x + v.w;
        __syncwarp();                               // every lane has read the slot before it is refilled
        const int64_t kk = warp + issued * n_warps;
        if (kk < n_access && lane == 0) issue(kk, slot);
        ++issued;
    }
    if (acc == 1.2345e30f) atomicAdd(sink, 1ull);
}

template <bool REWRITE>
__global__ void __launch_bounds__(BULK_WARPS * 32)
row_bulk_red_kernel(float4 *__restrict__ buf, uint64_t n_rows, int64_t n_access, uint64_t seed)
{
    __shared__ __align__(128) float4 s_row[BULK_WARPS][BULK_DEPTH][32];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int d = 0; d < BULK_DEPTH; ++d) s_row[wib][d][lane] = make_float4(1.f, 0.f, 0.f, -1.f);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    int64_t done = 0;
    for (int64_t k = warp; k < n_access; k += n_warps, ++done) {
        const int slot = (int)(done % BULK_DEPTH);
        if (REWRITE) {
            // the slot's previous reduction has read its source (at most DEPTH - 1 groups still pending)
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(BULK_DEPTH - 1) : "memory");
            __syncwarp();
            s_row[wib][slot][lane] = make_float4(1.f, (float)(k & 1), 0.f, -1.f);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
        }
        if (lane == 0) {
            const uint64_t row = mix64(seed + (uint64_t)k) % n_rows;
            asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], 512;"
                         ::"l"(buf + row * 32), "r"(smem_u32(&s_row[wib][slot][0])) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (!REWRITE) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(BULK_DEPTH - 1) : "memory");
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

__global__ void __launch_bounds__(256)
gather_sector16_kernel(const uint4 *__restrict__ buf, uint64_t n_sectors, int64_t n_access, uint64_t seed,
                       unsigned long long *__restrict__ sink)
{
    const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    uint32_t acc = 0;
    for (int64_t i = tid; i < n_access; i += 16 * stride) {
        uint4 v[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) {      // 16 independent sectors in flight per thread
            const int64_t j = i + u * stride;
            const uint64_t sct = mix64(seed + (uint64_t)j) % n_sectors;
            v[u] = j < n_access ? __ldg(buf + sct * 2) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < 16; ++u) acc += v[u].x ^ v[u].w;
    }
    if (acc == 0x12345678u) atomicAdd(sink, 1ull);
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
    } else if (mode == 5) {
        N2V_REQUIRE(sink, "sink is NULL");
        row_bulk_read_kernel<<<blocks, BULK_WARPS * 32, 0, stream>>>((const float4 *)buf, n_bytes / 512, n_access, seed, sink);
    } else if (mode == 6) {
        row_bulk_red_kernel<false><<<blocks, BULK_WARPS * 32, 0, stream>>>((float4 *)buf, n_bytes / 512, n_access, seed);
    } else if (mode == 7) {
        row_bulk_red_kernel<true><<<blocks, BULK_WARPS * 32, 0, stream>>>((float4 *)buf, n_bytes / 512, n_access, seed);
    } else if (mode == 8) {
        N2V_REQUIRE(sink, "sink is NULL");
        gather_sector16_kernel<<<blocks, 256, 0, stream>>>((const uint4 *)buf, n_bytes / 32, n_access, seed, sink);
    } else {
        set_error("n2v_random_gather_bench: unknown mode %d", mode);
        return N2V_EINVAL;
    }
    N2V_LAUNCH_CHECK();
    return N2V_OK;
}
