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

// ---- parser: walk file -> integer tokens + sentence offsets (LineSentence over a walk file) --------
namespace n2v {

__device__ __forceinline__ bool is_space(unsigned char c) { return c == ' ' || c == '\n' || c == '\t' || c == '\r' || c == '\v' || c == '\f'; }

// flags: token start (non-space byte whose predecessor is a space / start of text), line start
// (byte 0 and every byte after a '\n')
__global__ void parse_flags_kernel(const unsigned char *__restrict__ text, int64_t n, int32_t *__restrict__ tok_flag,
                                   int32_t *__restrict__ line_flag)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i > n) return;
    if (i == n) { tok_flag[i] = 0; line_flag[i] = 0; return; }
    const unsigned char c = text[i];
    const unsigned char p = i ? text[i - 1] : (unsigned char)'\n';
    tok_flag[i] = (!is_space(c) && is_space(p)) ? 1 : 0;
    line_flag[i] = (p == '\n') ? 1 : 0;
}

// every token start parses its digits; every line start records how many tokens precede it
__global__ void parse_fill_kernel(const unsigned char *__restrict__ text, int64_t n, const int32_t *__restrict__ tok_flag,
                                  const int32_t *__restrict__ line_flag, const int64_t *__restrict__ tok_idx,
                                  const int64_t *__restrict__ line_idx, long long *__restrict__ labels,
                                  int64_t *__restrict__ sent_off, int *__restrict__ bad)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (line_flag[i]) sent_off[line_idx[i]] = tok_idx[i];
    if (tok_flag[i]) {
        int64_t p = i;
        bool neg = false;
        if (text[p] == '-') { neg = true; ++p; }
        unsigned long long v = 0;
        int digits = 0;
        while (p < n && !is_space(text[p])) {
            const unsigned char c = text[p];
            if (c < '0' || c > '9' || digits >= 19) { *bad = 1; break; }
            v = v * 10ull + (unsigned long long)(c - '0');
            ++digits; ++p;
        }
        if (digits == 0) *bad = 1;
        labels[tok_idx[i]] = neg ? -(long long)v : (long long)v;
    }
}

}  // namespace n2v

extern "C" size_t n2v_parse_workspace_bytes(int64_t n_bytes)
{
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, (int32_t *)nullptr, (int64_t *)nullptr, n_bytes + 1);
    return tb + 256;
}

extern "C" int n2v_parse_walks_index(const unsigned char *text, int64_t n_bytes, int32_t *tok_flag, int32_t *line_flag,
                                     int64_t *tok_idx, int64_t *line_idx, void *workspace, size_t workspace_bytes,
                                     void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    N2V_REQUIRE(n_bytes >= 0, "negative size");
    N2V_REQUIRE(tok_flag && line_flag && tok_idx && line_idx && workspace && (n_bytes == 0 || text), "NULL buffer");
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, (int32_t *)nullptr, (int64_t *)nullptr, n_bytes + 1);
    if (tb > workspace_bytes) { set_error("parse workspace too small: need %zu", tb); return N2V_ENOMEM; }
    parse_flags_kernel<<<(unsigned)((n_bytes + 1 + 255) / 256), 256, 0, stream>>>(text, n_bytes, tok_flag, line_flag);
    N2V_LAUNCH_CHECK();
    size_t t1 = tb;
    N2V_CHECK_CUDA(cub::DeviceScan::ExclusiveSum(workspace, t1, tok_flag, tok_idx, n_bytes + 1, stream));
    t1 = tb;
    N2V_CHECK_CUDA(cub::DeviceScan::ExclusiveSum(workspace, t1, line_flag, line_idx, n_bytes + 1, stream));
    return N2V_OK;
}

extern "C" int n2v_parse_walks_fill(const unsigned char *text, int64_t n_bytes, const int32_t *tok_flag,
                                    const int32_t *line_flag, const int64_t *tok_idx, const int64_t *line_idx,
                                    int64_t *labels, int64_t *sent_off, int *bad_flag, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    N2V_REQUIRE(n_bytes >= 0, "negative size");
    if (n_bytes == 0) return N2V_OK;
    N2V_REQUIRE(text && tok_flag && line_flag && tok_idx && line_idx && labels && sent_off && bad_flag, "NULL buffer");
    parse_fill_kernel<<<(unsigned)((n_bytes + 255) / 256), 256, 0, stream>>>(
        text, n_bytes, tok_flag, line_flag, tok_idx, line_idx, (long long *)labels, sent_off, bad_flag);
    N2V_LAUNCH_CHECK();
    return N2V_OK;
}
