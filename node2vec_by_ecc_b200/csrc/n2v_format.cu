// Device-side walk-file formatter: the text corpus main_link.py writes between the
// walk and SGNS phases (src/main_link.py:237-239, :544-546 -- " ".join(map(str, walk)) per line) and
// LineSentence (:340) reads back. Pure byte work: token -> decimal digits + separator, HBM-bound
// (4-12 B read, ~7 B written per token).
#include <cub/cub.cuh>

#include "n2v_common.cuh"

namespace n2v {

__device__ __forceinline__ int dec_len(long long v)
{
    unsigned long long u = v < 0 ? (unsigned long long)(-(v + 1)) + 1ull : (unsigned long long)v;
    int n = v < 0 ? 2 : 1;
    while (u >= 10ull) { u /= 10ull; ++n; }
    return n;
}

__device__ __forceinline__ long long token_label(const int32_t *walks, const long long *labels, int64_t t)
{
    const int32_t id = walks[t];
    return labels ? labels[id] : (long long)id;
}

// bytes of every token slot: digits + 1 separator; slots past the walk's length take 0 bytes
// (an empty walk still yields its newline)
__global__ void format_sizes_kernel(const int32_t *__restrict__ walks, const int32_t *__restrict__ lens,
                                    int64_t n_walks, int32_t L, const long long *__restrict__ labels,
                                    int64_t *__restrict__ sizes)
{
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t n = n_walks * (int64_t)L;
    if (t > n) return;
    if (t == n) { sizes[t] = 0; return; }
    const int64_t w = t / L; const int32_t s = (int32_t)(t - w * L);
    const int32_t len = lens[w];
    int64_t b = 0;
    if (s < len) b = dec_len(token_label(walks, labels, t)) + 1;
    else if (s == 0) b = 1;
    sizes[t] = b;
}

__global__ void format_write_kernel(const int32_t *__restrict__ walks, const int32_t *__restrict__ lens,
                                    int64_t n_walks, int32_t L, const long long *__restrict__ labels,
                                    const int64_t *__restrict__ off, unsigned char *__restrict__ out)
{
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= n_walks * (int64_t)L) return;
    const int64_t w = t / L; const int32_t s = (int32_t)(t - w * L);
    const int32_t len = lens[w];
    const int64_t o = off[t], nb = off[t + 1] - o;
    if (nb == 0) return;
    if (s >= len) { out[o] = '\n'; return; }
    long long v = token_label(walks, labels, t);
    unsigned long long u = v < 0 ? (unsigned long long)(-(v + 1)) + 1ull : (unsigned long long)v;
    int64_t p = o + nb - 1;
    out[p--] = (s == len - 1) ? '\n' : ' ';
    do { out[p--] = (unsigned char)('0' + (int)(u % 10ull)); u /= 10ull; } while (u);
    if (v < 0) out[p] = '-';
}

}  // namespace n2v

using namespace n2v;

extern "C" size_t n2v_format_workspace_bytes(int64_t n_walks, int32_t L)
{
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, (int64_t *)nullptr, (int64_t *)nullptr, n_walks * (int64_t)L + 1);
    return tb + 256;
}

extern "C" int n2v_format_walks_offsets(const int32_t *walks, const int32_t *lens, int64_t n_walks, int32_t L,
                                        const int64_t *labels, int64_t *tok_off, void *workspace,
                                        size_t workspace_bytes, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    N2V_REQUIRE(n_walks >= 0 && L > 0, "bad size");
    N2V_REQUIRE(tok_off && workspace && (n_walks == 0 || (walks && lens)), "NULL buffer");
    const int64_t n = n_walks * (int64_t)L;
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, (int64_t *)nullptr, (int64_t *)nullptr, n + 1);
    if (tb > workspace_bytes) { set_error("format workspace too small: need %zu", tb); return N2V_ENOMEM; }
    format_sizes_kernel<<<(unsigned)((n + 1 + 255) / 256), 256, 0, stream>>>(walks, lens, n_walks, L,
                                                                             (const long long *)labels, tok_off);
    N2V_LAUNCH_CHECK();
    N2V_CHECK_CUDA(cub::DeviceScan::ExclusiveSum(workspace, tb, tok_off, tok_off, n + 1, stream));
    return N2V_OK;
}

extern "C" int n2v_format_walks_write(const int32_t *walks, const int32_t *lens, int64_t n_walks, int32_t L,
                                      const int64_t *labels, const int64_t *tok_off, unsigned char *out,
                                      void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    N2V_REQUIRE(n_walks >= 0 && L > 0, "bad size");
    if (n_walks == 0) return N2V_OK;
    N2V_REQUIRE(walks && lens && tok_off && out, "NULL buffer");
    const int64_t n = n_walks * (int64_t)L;
    format_write_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(walks, lens, n_walks, L,
                                                                         (const long long *)labels, tok_off, out);
    N2V_LAUNCH_CHECK();
    return N2V_OK;
}
