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

// ---- all-pairs similarity without the matrix ------------------------------------------------------
// link_prediction / make_links_and_score / links_score (src/main_link.py:70-171) score EVERY candidate
// pair (user x item, or every unordered pair of nodes) and keep the best k; build_user_sim_matrx and
// get_add_edge_by_* (:368-453) score every user x user pair and keep, per user, the pairs above a
// threshold or the top share. Both are S = normalise(A) . normalise(B)^T followed by a selection. The
// kernel below computes S tile by tile on the fp32 pipes (64 x 64 outputs per block, K in chunks of 32
// through shared memory, 4 x 4 outputs per thread -- fp32 FMA so the scores are the reference's float32
// cosines, not tf32 ones) and applies the selection in the epilogue: a pair is EMITTED iff its score is
// above thr_row[row] (or the scalar thr) and it survives the masks; S itself is never written.
// Selections that need an order (global / per-row top-k) choose the threshold from a sample
// (n2v_cosine_pairs) and sort the few emitted candidates: node2vec_by_ecc_b200/scoring.py.
constexpr int ST_TILE = 64, ST_K = 32, ST_THREADS = 256;

__global__ void row_center_norm_kernel(const float *__restrict__ emb, int32_t dim, const int32_t *__restrict__ rows,
                                       int64_t n, int centered, float *__restrict__ mean, float *__restrict__ inv_norm)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (warp >= n) return;
    const int32_t r = rows[warp];
    float sum = 0.f, sq = 0.f;
    if (r >= 0)
        for (int c = lane; c < dim; c += 32) { const float v = emb[(int64_t)r * dim + c]; sum += v; sq += v * v; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o); sq += __shfl_xor_sync(0xFFFFFFFFu, sq, o); }
    if (lane == 0) {
        const float m = centered ? sum / (float)dim : 0.f;         // pearsonr == cosine of the centred rows
        const float ss = sq - (float)dim * m * m;
        mean[warp] = m;
        inv_norm[warp] = ss > 0.f ? rsqrtf(ss) : 0.f;
    }
}

struct SimArgs {
    const float *emb; int32_t dim;
    const int32_t *rows_a, *rows_b; int32_t n_a, n_b;
    const float *mean_a, *inv_a, *mean_b, *inv_b;
    const float *thr_row; float thr;
    int32_t upper_only, skip_diagonal;
    const long long *exclude; int64_t n_exclude;          // sorted keys a_pos * n_b + b_pos
    int32_t *out_a, *out_b; float *out_s; int64_t capacity;
    unsigned long long *count;
};

__global__ void __launch_bounds__(ST_THREADS)
sim_threshold_kernel(SimArgs g)
{
    __shared__ float sA[ST_K][ST_TILE + 4], sB[ST_K][ST_TILE + 4];
    __shared__ unsigned long long s_excl[ST_TILE];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int32_t r0 = blockIdx.y * ST_TILE, c0 = blockIdx.x * ST_TILE;
    if (g.upper_only && c0 + ST_TILE - 1 <= r0) return;            // tile entirely on or below the diagonal
    // accumulators as packed fp32 pairs (FFMA2: one issue slot for two fused multiply-adds): acc2[i][h] holds
    // outputs (row i, columns 2h and 2h + 1)
    unsigned long long acc2[4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i) acc2[i][0] = acc2[i][1] = 0ull;
    // the tile's excluded pairs: one row per thread, a bit per column
    if (tid < ST_TILE) {
        unsigned long long m = 0ull;
        const int32_t r = r0 + tid;
        if (g.n_exclude > 0 && r < g.n_a) {
            const long long lo_key = (long long)r * g.n_b + c0, hi_key = lo_key + ST_TILE;
            int64_t lo = 0, hi = g.n_exclude;
            while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (g.exclude[mid] < lo_key) lo = mid + 1; else hi = mid; }
            for (; lo < g.n_exclude && g.exclude[lo] < hi_key; ++lo) m |= 1ull << (int)(g.exclude[lo] - lo_key);
        }
        s_excl[tid] = m;
    }
    for (int32_t k0 = 0; k0 < g.dim; k0 += ST_K) {
        // stage 64 rows x 32 floats of each side, K-major in shared memory (centred when asked)
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            const int idx = tid + it * ST_THREADS;                  // 512 float4 per side
            const int row = idx >> 3, c4 = (idx & 7) * 4;
            float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
            if (k0 + c4 < g.dim) {
                const int32_t ra = r0 + row < g.n_a ? g.rows_a[r0 + row] : -1;
                const int32_t rb = c0 + row < g.n_b ? g.rows_b[c0 + row] : -1;
                if (ra >= 0) {
                    va = __ldg(reinterpret_cast<const float4 *>(g.emb + (int64_t)ra * g.dim + k0 + c4));
                    const float m = g.mean_a[r0 + row];
                    va.x -= m; va.y -= m; va.z -= m; va.w -= m;
                }
                if (rb >= 0) {
                    vb = __ldg(reinterpret_cast<const float4 *>(g.emb + (int64_t)rb * g.dim + k0 + c4));
                    const float m = g.mean_b[c0 + row];
                    vb.x -= m; vb.y -= m; vb.z -= m; vb.w -= m;
                }
            }
            sA[c4][row] = va.x; sA[c4 + 1][row] = va.y; sA[c4 + 2][row] = va.z; sA[c4 + 3][row] = va.w;
            sB[c4][row] = vb.x; sB[c4 + 1][row] = vb.y; sB[c4 + 2][row] = vb.z; sB[c4 + 3][row] = vb.w;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < ST_K; ++kk) {
            const float4 a4 = *reinterpret_cast<const float4 *>(&sA[kk][ty * 4]);
            const float4 b4 = *reinterpret_cast<const float4 *>(&sB[kk][tx * 4]);
            const float av[4] = {a4.x, a4.y, a4.z, a4.w};
            unsigned long long b01, b23;
            asm("mov.b64 %0, {%1, %2};" : "=l"(b01) : "f"(b4.x), "f"(b4.y));
            asm("mov.b64 %0, {%1, %2};" : "=l"(b23) : "f"(b4.z), "f"(b4.w));
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                unsigned long long aa;
                asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(av[i]));
                asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc2[i][0]) : "l"(aa), "l"(b01));
                asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc2[i][1]) : "l"(aa), "l"(b23));
            }
        }
        __syncthreads();
    }
    // epilogue: normalise, mask, select, emit
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        asm("mov.b64 {%0, %1}, %2;" : "=f"(acc[i][0]), "=f"(acc[i][1]) : "l"(acc2[i][0]));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(acc[i][2]), "=f"(acc[i][3]) : "l"(acc2[i][1]));
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int32_t r = r0 + ty * 4 + i;
        if (r >= g.n_a) continue;
        const float ia = g.inv_a[r];
        const float thr = g.thr_row ? g.thr_row[r] : g.thr;
        const unsigned long long ex = s_excl[ty * 4 + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int32_t c = c0 + tx * 4 + j;
            if (c >= g.n_b) continue;
            if (g.upper_only && c <= r) continue;
            const float sc = (g.skip_diagonal && c == r) ? 0.f : acc[i][j] * ia * g.inv_b[c];
            if (!(sc > thr) || ((ex >> (tx * 4 + j)) & 1ull)) continue;
            const unsigned long long o = atomicAdd(g.count, 1ull);
            if ((int64_t)o < g.capacity) { g.out_a[o] = r; g.out_b[o] = c; g.out_s[o] = sc; }
        }
    }
}

}  // namespace n2v

using namespace n2v;

extern "C" int n2v_row_norms(const float *emb, int32_t dim, const int32_t *rows, int64_t n, int centered,
                             float *mean, float *inv_norm, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    N2V_REQUIRE(n >= 0 && dim > 0, "bad size");
    if (n == 0) return N2V_OK;
    N2V_REQUIRE(emb && rows && mean && inv_norm, "NULL buffer");
    row_center_norm_kernel<<<(unsigned)((n + 7) / 8), 256, 0, stream>>>(emb, dim, rows, n, centered, mean, inv_norm);
    N2V_LAUNCH_CHECK();
    return N2V_OK;
}

extern "C" int n2v_sim_threshold(const float *emb, int32_t dim, const int32_t *rows_a, int32_t n_a,
                                 const int32_t *rows_b, int32_t n_b, const float *mean_a, const float *inv_a,
                                 const float *mean_b, const float *inv_b, const float *thr_row, float thr,
                                 int upper_only, int skip_diagonal, const long long *exclude, int64_t n_exclude,
                                 int32_t *out_a, int32_t *out_b, float *out_score, int64_t capacity,
                                 unsigned long long *count, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    N2V_REQUIRE(n_a >= 0 && n_b >= 0 && dim > 0 && dim % 4 == 0, "dim must be a positive multiple of 4");
    N2V_REQUIRE(capacity >= 0 && n_exclude >= 0, "negative size");
    if (n_a == 0 || n_b == 0) return N2V_OK;
    N2V_REQUIRE(emb && rows_a && rows_b && mean_a && inv_a && mean_b && inv_b && count, "NULL buffer");
    N2V_REQUIRE(capacity == 0 || (out_a && out_b && out_score), "NULL output buffer");
    N2V_REQUIRE(n_exclude == 0 || exclude, "exclude is NULL");
    SimArgs g{emb, dim, rows_a, rows_b, n_a, n_b, mean_a, inv_a, mean_b, inv_b, thr_row, thr, upper_only, skip_diagonal,
              exclude, n_exclude, out_a, out_b, out_score, capacity, count};
    const dim3 grid((unsigned)((n_b + ST_TILE - 1) / ST_TILE), (unsigned)((n_a + ST_TILE - 1) / ST_TILE));
    N2V_REQUIRE(grid.y < 65536u, "too many rows for one launch (<= 4,194,240): call per block of rows");
    sim_threshold_kernel<<<grid, ST_THREADS, 0, stream>>>(g);
    N2V_LAUNCH_CHECK();
    return N2V_OK;
}

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
