// Error plumbing, device queries and the device CSR builder (north-star subsystem 1).
#include <cub/cub.cuh>
#include <stdarg.h>
#include <string.h>

#include "n2v_common.cuh"

namespace n2v {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int sm_count()
{
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    return n;
}

static inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// ---- CSR builder ---------------------------------------------------------------------------
// key = src<<32 | dst, value = input sequence number (forward arc of edge i: 2i, reverse: 2i+1,
// so that "last in input order wins" == "largest sequence number of a key run wins").
__global__ void csr_make_keys(const int32_t *__restrict__ src, const int32_t *__restrict__ dst,
                              int64_t m, int undirected, uint64_t *__restrict__ keys,
                              uint32_t *__restrict__ seq)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= m) return;
    uint32_t s = (uint32_t)src[i], d = (uint32_t)dst[i];
    if (undirected) {
        keys[2 * i] = ((uint64_t)s << 32) | d;      seq[2 * i] = (uint32_t)(2 * i);
        keys[2 * i + 1] = ((uint64_t)d << 32) | s;  seq[2 * i + 1] = (uint32_t)(2 * i + 1);
    } else {
        keys[i] = ((uint64_t)s << 32) | d;          seq[i] = (uint32_t)i;
    }
}

// flag[i] = 1 when sorted position i is the last of its key run (the stable sort keeps the
// sequence numbers ascending inside a run)
__global__ void csr_flag_last(const uint64_t *__restrict__ keys, int64_t M, int32_t *__restrict__ flag)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= M) return;
    flag[i] = (i + 1 == M || keys[i] != keys[i + 1]) ? 1 : 0;
}

__global__ void csr_scatter(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ seq,
                            const int32_t *__restrict__ flag, const int64_t *__restrict__ pos,
                            const double *__restrict__ w, int undirected, int64_t M,
                            int32_t *__restrict__ col, double *__restrict__ w_out,
                            unsigned long long *__restrict__ deg, int64_t *__restrict__ nnz_out)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= M) return;
    if (flag[i]) {
        int64_t o = pos[i];
        uint64_t k = keys[i];
        col[o] = (int32_t)(uint32_t)k;
        if (w_out) w_out[o] = w ? w[undirected ? (seq[i] >> 1) : seq[i]] : 1.0;
        atomicAdd(deg + (k >> 32) + 1, 1ull);
        if (i + 1 == M) *nnz_out = o + 1;
    }
}

struct CsrWs {
    uint64_t *keys_a, *keys_b;
    uint32_t *seq_a, *seq_b;
    int32_t *flag;
    int64_t *pos;
    void *cub_tmp;
    size_t cub_bytes, total;
};

static CsrWs carve_csr_ws(void *base, int64_t M, int32_t n_nodes)
{
    CsrWs ws;
    size_t sort_bytes = 0, scan_bytes = 0, scan2_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (uint64_t *)nullptr, (uint64_t *)nullptr,
                                    (uint32_t *)nullptr, (uint32_t *)nullptr, M);
    cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, (int32_t *)nullptr, (int64_t *)nullptr, M);
    cub::DeviceScan::InclusiveSum(nullptr, scan2_bytes, (int64_t *)nullptr, (int64_t *)nullptr,
                                  (int64_t)n_nodes + 1);
    ws.cub_bytes = align_up(sort_bytes > scan_bytes ? (sort_bytes > scan2_bytes ? sort_bytes : scan2_bytes)
                                                    : (scan_bytes > scan2_bytes ? scan_bytes : scan2_bytes));
    char *p = (char *)base;
    size_t off = 0;
    auto take = [&](size_t bytes) { char *r = p ? p + off : nullptr; off += align_up(bytes); return r; };
    ws.keys_a = (uint64_t *)take(sizeof(uint64_t) * (size_t)M);
    ws.keys_b = (uint64_t *)take(sizeof(uint64_t) * (size_t)M);
    ws.seq_a = (uint32_t *)take(sizeof(uint32_t) * (size_t)M);
    ws.seq_b = (uint32_t *)take(sizeof(uint32_t) * (size_t)M);
    ws.flag = (int32_t *)take(sizeof(int32_t) * (size_t)M);
    ws.pos = (int64_t *)take(sizeof(int64_t) * (size_t)M);
    ws.cub_tmp = take(ws.cub_bytes);
    ws.total = off;
    return ws;
}

}  // namespace n2v

using namespace n2v;

extern "C" const char *n2v_last_error(void) { return g_err; }
extern "C" int n2v_version(void) { return 100; }
extern "C" int n2v_sm_count(void)
{
    int n = sm_count();
    if (n < 0) { set_error("no CUDA device"); return N2V_ECUDA; }
    return n;
}

extern "C" size_t n2v_csr_workspace_bytes(int64_t m, int32_t n_nodes, int undirected)
{
    int64_t M = undirected ? 2 * m : m;
    if (M < 1) M = 1;
    return carve_csr_ws(nullptr, M, n_nodes).total;
}

extern "C" int n2v_csr_from_coo(const int32_t *src, const int32_t *dst, const double *w, int64_t m,
                                int32_t n_nodes, int undirected, void *workspace,
                                size_t workspace_bytes, int64_t *row_ptr, int32_t *col,
                                double *w_out, int64_t *nnz_out, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    N2V_REQUIRE(n_nodes >= 0 && m >= 0, "negative size");
    N2V_REQUIRE(row_ptr && nnz_out, "row_ptr / nnz_out is NULL");
    const int64_t M = undirected ? 2 * m : m;
    N2V_REQUIRE(M < (int64_t)0xFFFFFFFFll, "more than 2^32-1 arcs");
    N2V_CHECK_CUDA(cudaMemsetAsync(row_ptr, 0, sizeof(int64_t) * ((size_t)n_nodes + 1), stream));
    N2V_CHECK_CUDA(cudaMemsetAsync(nnz_out, 0, sizeof(int64_t), stream));
    if (M == 0) return N2V_OK;
    N2V_REQUIRE(src && dst && col && workspace, "NULL buffer");
    CsrWs ws = carve_csr_ws(workspace, M, n_nodes);
    if (ws.total > workspace_bytes) { set_error("csr workspace too small: need %zu", ws.total); return N2V_ENOMEM; }
    const int T = 256;
    csr_make_keys<<<(unsigned)((m + T - 1) / T), T, 0, stream>>>(src, dst, m, undirected, ws.keys_a, ws.seq_a);
    N2V_LAUNCH_CHECK();
    size_t tb = ws.cub_bytes;
    int end_bit = 32;   // dst bits + as many src bits as n_nodes needs
    for (int64_t v = n_nodes; v > 0; v >>= 1) ++end_bit;
    if (end_bit > 64) end_bit = 64;
    N2V_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(ws.cub_tmp, tb, ws.keys_a, ws.keys_b, ws.seq_a,
                                                   ws.seq_b, M, 0, end_bit, stream));
    csr_flag_last<<<(unsigned)((M + T - 1) / T), T, 0, stream>>>(ws.keys_b, M, ws.flag);
    N2V_LAUNCH_CHECK();
    tb = ws.cub_bytes;
    N2V_CHECK_CUDA(cub::DeviceScan::ExclusiveSum(ws.cub_tmp, tb, ws.flag, ws.pos, M, stream));
    csr_scatter<<<(unsigned)((M + T - 1) / T), T, 0, stream>>>(
        ws.keys_b, ws.seq_b, ws.flag, ws.pos, w, undirected, M, col, w_out,
        (unsigned long long *)row_ptr, nnz_out);
    N2V_LAUNCH_CHECK();
    tb = ws.cub_bytes;
    N2V_CHECK_CUDA(cub::DeviceScan::InclusiveSum(ws.cub_tmp, tb, row_ptr, row_ptr, (int64_t)n_nodes + 1, stream));
    return N2V_OK;
}
