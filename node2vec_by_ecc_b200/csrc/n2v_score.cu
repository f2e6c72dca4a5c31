// Link scoring on device: cosine of embedding rows for a batch of node pairs --
// link_score 'cos' (src/main_link.py:43-49) as looped by get_roc_score (:173-189).
// One warp per pair, float4 loads, HBM-bound (2 rows of d*4 bytes per pair).
#include "n2v_common.cuh"

namespace n2v {

__global__ void __launch_bounds__(256)
cosine_pairs_kernel(const float *__restrict__ emb, int32_t dim, const int32_t *__restrict__ a,
                    const int32_t *__restrict__ b, int64_t n_pairs, float *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp; i < n_pairs; i += n_warps) {
        const int32_t ia = a[i], ib = b[i];
        float dot = 0.f, na = 0.f, nb = 0.f;
        if (ia >= 0 && ib >= 0) {
            const float4 *ra = reinterpret_cast<const float4 *>(emb + (int64_t)ia * dim);
            const float4 *rb = reinterpret_cast<const float4 *>(emb + (int64_t)ib * dim);
            for (int c = lane; c * 4 < dim; c += 32) {
                const float4 x = __ldg(ra + c), y = __ldg(rb + c);
                dot += x.x * y.x + x.y * y.y + x.z * y.z + x.w * y.w;
                na += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
                nb += y.x * y.x + y.y * y.y + y.z * y.z + y.w * y.w;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            dot += __shfl_xor_sync(0xFFFFFFFFu, dot, o);
            na += __shfl_xor_sync(0xFFFFFFFFu, na, o);
            nb += __shfl_xor_sync(0xFFFFFFFFu, nb, o);
        }
        // a word missing from the vocabulary scores 0, as link_score's except branch (:47-48)
        if (lane == 0) out[i] = (na > 0.f && nb > 0.f) ? dot * rsqrtf(na) * rsqrtf(nb) : 0.f;
    }
}

}  // namespace n2v

using namespace n2v;

extern "C" int n2v_cosine_pairs(const float *emb, int32_t dim, const int32_t *a, const int32_t *b,
                                int64_t n_pairs, float *out, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    N2V_REQUIRE(n_pairs >= 0 && dim > 0 && dim % 4 == 0, "dim must be a positive multiple of 4");
    if (n_pairs == 0) return N2V_OK;
    N2V_REQUIRE(emb && a && b && out, "NULL buffer");
    int sms = sm_count();
    if (sms <= 0) { set_error("no CUDA device"); return N2V_ECUDA; }
    int64_t blocks = (n_pairs + 7) / 8;
    if (blocks > (int64_t)sms * 8) blocks = (int64_t)sms * 8;
    cosine_pairs_kernel<<<(unsigned)blocks, 256, 0, stream>>>(emb, dim, a, b, n_pairs, out);
    N2V_LAUNCH_CHECK();
    return N2V_OK;
}
