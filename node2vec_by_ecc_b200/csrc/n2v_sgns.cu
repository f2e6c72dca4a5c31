// Skip-gram negative sampling (north-star subsystem 4): the arithmetic of gensim 3.2.0
// word2vec_inner.pyx train_batch_sg / fast_sentence_sg_neg, as called through
// learn_embeddings (src/main.py:82-90).
//
// Work decomposition: one warp owns one sentence (walk) at a time and runs it exactly like one
// gensim worker thread runs a sentence: sub-sample, shrink the window per position, then for
// every (centre i, context j) pair: input row syn0[w_j], targets syn1neg[w_i] (label 1) and
// `negative` draws from the count^0.75 table (label 0, draws equal to w_i skipped), sigmoid from
// the 1000-entry table, immediate updates. Warps run concurrently against the same tables with
// no locks -- gensim's Hogwild, with `grid_warps` as the Hogwild width.
//
// Data movement per pair (d=128, K=5): 7 rows read + 7 rows written, 512 B each = 7,168 B; each
// lane holds one float4 of every row (16-byte vector loads, a full 512-byte row per warp
// instruction), dots are reduced with 5 xor-shuffles. The kernel is HBM random-row bound
// (0.64 flop/B); tensor cores do not apply.
#include <cub/cub.cuh>
#include <type_traits>

#include "n2v_common.cuh"
#include "n2v_sgns_stage.cuh"

namespace n2v {

// ---- vocabulary ---------------------------------------------------------------------------------
__global__ void vocab_count_kernel(const int32_t *__restrict__ tokens, int64_t n, int32_t n_ids,
                                   unsigned long long *__restrict__ counts)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        int32_t t = tokens[i];
        if (t >= 0 && t < n_ids) atomicAdd(counts + t, 1ull);
    }
}

__global__ void sgns_pow_kernel(const unsigned long long *__restrict__ counts, int32_t V,
                                double *__restrict__ powed, double *__restrict__ countd)
{
    int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V) return;
    double c = (double)counts[i];
    powed[i] = pow(c, 0.75);
    countd[i] = c;
}

__global__ void sgns_prepare_kernel(const unsigned long long *__restrict__ counts, int32_t V,
                                    double sample, const double *__restrict__ cum_pow,
                                    const double *__restrict__ cum_cnt,
                                    uint32_t *__restrict__ keep_thr, uint32_t *__restrict__ cum_table)
{
    int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V) return;
    const double retain_total = cum_cnt[V - 1], total_pow = cum_pow[V - 1];
    double threshold;
    if (sample <= 0.0) threshold = retain_total;
    else if (sample < 1.0) threshold = sample * retain_total;
    else threshold = floor(sample * (3.0 + sqrt(5.0)) / 2.0);
    const double v = (double)counts[i];
    double prob = (sqrt(v / threshold) + 1.0) * (threshold / v);   // scale_vocab
    if (!(prob < 1.0)) prob = 1.0;
    double si = round(prob * 4294967296.0);                        // sample_int
    keep_thr[i] = si >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)si; // dropped iff sample_int < r32
    cum_table[i] = (uint32_t)round(cum_pow[i] / total_pow * 2147483647.0);   // make_cum_table
}

// bucket_lo[b] = bisect_left(cum_table, b << shift): the search for a draw r then only spans
// [bucket_lo[r >> shift], bucket_lo[(r >> shift) + 1]].
__global__ void sgns_bucket_kernel(const uint32_t *__restrict__ cum_table, int32_t V,
                                   int32_t *__restrict__ bucket_lo, int32_t bucket_bits)
{
    const int32_t nb = (1 << bucket_bits) + 1;
    int32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    const uint64_t r = (uint64_t)b << (31 - bucket_bits);
    int32_t lo = 0, hi = V;
    while (lo < hi) { int32_t mid = (lo + hi) >> 1; if ((uint64_t)cum_table[mid] < r) lo = mid + 1; else hi = mid; }
    bucket_lo[b] = lo;
}

// rows of this table: local row l holds vocabulary row l * n_parts + part (n_parts = 1: the whole table)
__global__ void sgns_init_kernel(float *__restrict__ syn0, float *__restrict__ syn1neg, int32_t n_local,
                                 int32_t dim, uint32_t k0, uint32_t k1, int32_t part, int32_t n_parts)
{
    const int32_t chunks = (dim + 3) >> 2;
    int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= (int64_t)n_local * chunks) return;
    const int32_t row = (int32_t)(t / chunks), c = (int32_t)(t % chunks);
    const Philox4 r = philox4x32_10((uint32_t)(row * n_parts + part), (uint32_t)c, 0x53594E30u, 0u, k0, k1);
    const uint32_t rr[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int32_t j = c * 4 + k;
        if (j < dim) {
            syn0[(int64_t)row * dim + j] = (float)(((double)rr[k] * (1.0 / 4294967296.0) - 0.5) / (double)dim);
            if (syn1neg) syn1neg[(int64_t)row * dim + j] = 0.0f;
        }
    }
}

// (train_pair / apply_target / RowsFlat / RowsSharded / warp_sum: n2v_sgns_stage.cuh)

// lane n (< negative) draws negative n of pair (i, j)
__device__ __forceinline__ int32_t draw_pair_negatives(const SgnsArgs &a, const WarpSentence &ws, int32_t i,
                                                       int32_t j, uint64_t gs, uint32_t ep8, uint32_t k0,
                                                       uint32_t k1, int lane)
{
    int32_t my_t = -1;
    if (lane < a.p.negative) {
        const Philox4 r = philox4x32_10((uint32_t)gs, (uint32_t)(gs >> 32),
                                        ((uint32_t)ws.pos[i] << 16) | (uint32_t)ws.pos[j],
                                        ep8 | (uint32_t)(1 + (lane >> 2)), k0, k1);
        const uint32_t rr = (lane & 3) == 0 ? r.x : (lane & 3) == 1 ? r.y : (lane & 3) == 2 ? r.z : r.w;
        my_t = draw_negative(rr, a.cum_table, a.bucket_lo, a.p.V, a.p.bucket_bits);
    }
    return my_t;
}

// ---- v1: generic kernel (any dim <= 1024, any negative <= 16), one pair at a time ------------------
// NV = float4 chunks per lane (dim <= 128*NV). ATOMIC selects red.global.add.v4.f32 updates.
template <int NV, bool ATOMIC>
__global__ void __launch_bounds__(SGNS_BLOCK)
sgns_train_kernel(SgnsArgs a)
{
    __shared__ int32_t s_idx[SGNS_BLOCK / 32][SGNS_SMEM_TOKENS];
    __shared__ uint16_t s_pos[SGNS_BLOCK / 32][SGNS_SMEM_TOKENS];
    __shared__ uint8_t s_rw[SGNS_BLOCK / 32][SGNS_SMEM_TOKENS];
    __shared__ float s_exp[EXP_TABLE_SIZE];
    for (int i = threadIdx.x; i < EXP_TABLE_SIZE; i += blockDim.x) s_exp[i] = exp_table_entry(i);
    __syncthreads();

    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const WarpSentence ws{s_idx[wib], s_pos[wib], s_rw[wib]};
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = a.p.grid_warps;
    if (warp >= n_warps) return;
    const int32_t dim = a.p.dim, window = a.p.window, negative = a.p.negative;
    const uint32_t k0 = (uint32_t)a.p.seed, k1 = (uint32_t)(a.p.seed >> 32);
    const uint32_t ep8 = a.p.epoch << 8;
    bool act[NV];
#pragma unroll
    for (int c = 0; c < NV; ++c) act[c] = (c * 128 + lane * 4) < dim;
    unsigned long long pairs = 0;

    for (int64_t s = warp; s < a.n_sent; s += n_warps) {
        const int64_t tb = a.sent_off ? a.sent_off[s] : s * (int64_t)a.stride;
        int64_t tl = a.sent_off ? a.sent_off[s + 1] - tb : (int64_t)a.stride;
        if (tl > a.p.max_sentence_len) tl = a.p.max_sentence_len;
        const uint64_t gs = (uint64_t)(a.sent_id_base + s);
        const float alpha = job_alpha(a.p, s);
        int64_t t_next = 0;
        int32_t n_kept = 0, c_lo = 0, c_hi = 0;
        bool first_chunk = true;
        while (next_chunk(a, ws, tb, tl, t_next, gs, ep8, k0, k1, lane, n_kept, c_lo, c_hi, first_chunk)) {
            for (int32_t i = c_lo; i < c_hi; ++i) {
                const int32_t centre = ws.idx[i];
                int32_t j = i - window + ws.rw[i]; if (j < 0) j = 0;
                int32_t kend = i + window + 1 - ws.rw[i]; if (kend > n_kept) kend = n_kept;
                for (; j < kend; ++j) {
                    if (j == i) continue;
                    const int32_t my_t = draw_pair_negatives(a, ws, i, j, gs, ep8, k0, k1, lane);
                    train_pair<NV, ATOMIC>(RowsFlat{a.syn0, a.syn1neg, dim}, dim, centre, ws.idx[j], my_t, negative, alpha, act, s_exp, lane);
                    ++pairs;
                }
            }
            __syncwarp();
        }
    }
    if (lane == 0 && a.pairs_out && pairs) atomicAdd(a.pairs_out, pairs);
}

// ---- v2: the reference configuration (dim <= 128, negative == 5) --------------------------------
// Same arithmetic in the same order as v1 / fast_sentence_sg_neg, restructured for latency:
//  * centre-major: syn1neg[centre] (the positive target of every pair of a centre) is read once,
//    carried in registers across the ~10 context pairs and written back as ONE reduction -- no other
//    pair of this warp can touch that row meanwhile (negatives equal to the centre are skipped);
//  * the negative draws of pair k+1 (Philox -> bucket index -> bisect, 3-4 dependent L2 reads) are
//    issued between the row loads of pair k and their first use, off the critical path;
//  * rows are loaded with ld.global.cg (no reuse inside an SM; RED updates happen at L2).
template <bool ATOMIC, int MINB>
__global__ void __launch_bounds__(SGNS_BLOCK, MINB)
sgns_train_kernel_v2(SgnsArgs a)
{
    constexpr int FN = 5;
    __shared__ int32_t s_idx[SGNS_BLOCK / 32][SGNS_SMEM_TOKENS];
    __shared__ uint16_t s_pos[SGNS_BLOCK / 32][SGNS_SMEM_TOKENS];
    __shared__ uint8_t s_rw[SGNS_BLOCK / 32][SGNS_SMEM_TOKENS];
    __shared__ float s_exp[EXP_TABLE_SIZE];
    for (int i = threadIdx.x; i < EXP_TABLE_SIZE; i += blockDim.x) s_exp[i] = exp_table_entry(i);
    __syncthreads();

    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const WarpSentence ws{s_idx[wib], s_pos[wib], s_rw[wib]};
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = a.p.grid_warps;
    if (warp >= n_warps) return;
    const int32_t dim = a.p.dim, window = a.p.window;
    const uint32_t k0 = (uint32_t)a.p.seed, k1 = (uint32_t)(a.p.seed >> 32);
    const uint32_t ep8 = a.p.epoch << 8;
    const bool on = lane * 4 < dim;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float *const syn0 = a.syn0, *const syn1neg = a.syn1neg;
    unsigned long long pairs = 0;

    for (int64_t s = warp; s < a.n_sent; s += n_warps) {
        const int64_t tb = a.sent_off ? a.sent_off[s] : s * (int64_t)a.stride;
        int64_t tl = a.sent_off ? a.sent_off[s + 1] - tb : (int64_t)a.stride;
        if (tl > a.p.max_sentence_len) tl = a.p.max_sentence_len;
        const uint64_t gs = (uint64_t)(a.sent_id_base + s);
        const float alpha = job_alpha(a.p, s);
        int64_t t_next = 0;
        int32_t n_kept = 0, c_lo = 0, c_hi = 0;
        bool first_chunk = true;
        while (next_chunk(a, ws, tb, tl, t_next, gs, ep8, k0, k1, lane, n_kept, c_lo, c_hi, first_chunk)) {
            // pair cursor (i, j) in gensim's order: centres ascending, contexts ascending, j != i
            int32_t i = c_lo - 1, j = 0, kend = 0;
            auto seek = [&]() -> bool {
                for (;;) {
                    if (i >= c_hi) return false;
                    if (j < kend) { if (j != i) return true; ++j; continue; }
                    if (++i >= c_hi) return false;
                    j = i - window + ws.rw[i]; if (j < 0) j = 0;
                    kend = i + window + 1 - ws.rw[i]; if (kend > n_kept) kend = n_kept;
                }
            };
            bool more = seek();
            int32_t t_cur = more ? draw_pair_negatives(a, ws, i, j, gs, ep8, k0, k1, lane) : -1;
            int32_t carried = -1;                        // vocabulary index of the carried centre row
            int32_t carried_i = -1;
            float4 pos_row = zero4, pos_delta = zero4;
            while (more) {
                const int32_t centre = ws.idx[i], ctx = ws.idx[j];
                if (i != carried_i) {                    // new centre: write the old row back, fetch the new one
                    if (carried >= 0) add_row<ATOMIC>(syn1neg + (int64_t)carried * dim, lane, pos_delta, pos_row, on);
                    carried = centre; carried_i = i;
                    pos_row = on ? ldcg4(syn1neg + (int64_t)centre * dim, lane) : zero4;
                    pos_delta = zero4;
                }
                // B: all rows of this pair in flight
                float *const row1p = syn0 + (int64_t)ctx * dim;
                const float4 row1 = on ? ldcg4(row1p, lane) : zero4;
                int32_t tg[FN];
                float4 r2[FN];
                bool dup = false;
#pragma unroll
                for (int d = 0; d < FN; ++d) tg[d] = __shfl_sync(0xFFFFFFFFu, t_cur, d);
#pragma unroll
                for (int d1 = 0; d1 < FN; ++d1)
#pragma unroll
                    for (int d2 = d1 + 1; d2 < FN; ++d2) dup |= (tg[d1] == tg[d2]);
#pragma unroll
                for (int d = 0; d < FN; ++d)
                    r2[d] = (on && tg[d] != centre) ? ldcg4(syn1neg + (int64_t)tg[d] * dim, lane) : zero4;
                // A (next pair): advance the cursor, draw its negatives while the rows arrive
                ++j;
                more = seek();
                const int32_t t_nxt = more ? draw_pair_negatives(a, ws, i, j, gs, ep8, k0, k1, lane) : -1;
                // C: fast_sentence_sg_neg, targets in order: centre (label 1), then the negatives
                float4 work = zero4;
                float f[FN + 1];
                f[0] = dot4(row1, pos_row);
#pragma unroll
                for (int d = 0; d < FN; ++d) f[d + 1] = dot4(row1, r2[d]);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                    for (int d = 0; d <= FN; ++d) f[d] += __shfl_xor_sync(0xFFFFFFFFu, f[d], o);
                if (f[0] > -(float)MAX_EXP && f[0] < (float)MAX_EXP) {
                    const float g = (1.0f - s_exp[(int)((f[0] + (float)MAX_EXP) * (float)(EXP_TABLE_SIZE / MAX_EXP / 2))]) * alpha;
                    axpy4(work, g, pos_row);                                  // work += g * syn1neg[centre]
                    axpy4(pos_row, g, row1); axpy4(pos_delta, g, row1);       // syn1neg[centre] += g * row1
                }
                if (!dup) {
#pragma unroll
                    for (int d = 0; d < FN; ++d) {
                        if (tg[d] == centre) continue;                        // skipped, not redrawn
                        if (f[d + 1] <= -(float)MAX_EXP || f[d + 1] >= (float)MAX_EXP) continue;
                        const float g = (0.0f - s_exp[(int)((f[d + 1] + (float)MAX_EXP) * (float)(EXP_TABLE_SIZE / MAX_EXP / 2))]) * alpha;
                        axpy4(work, g, r2[d]);
                        float4 upd = r2[d];
                        axpy4(upd, g, row1);
                        add_row<ATOMIC>(syn1neg + (int64_t)tg[d] * dim, lane,
                                        make_float4(g * row1.x, g * row1.y, g * row1.z, g * row1.w), upd, on);
                    }
                } else {
                    // a repeated negative must see the row as updated by its first occurrence
                    for (int d = 0; d < FN; ++d) {
                        const int32_t t = __shfl_sync(0xFFFFFFFFu, t_cur, d);
                        if (t == centre) continue;
                        float *const rp = syn1neg + (int64_t)t * dim;
                        const float4 row2 = on ? ldcg4(rp, lane) : zero4;
                        const float ff = warp_sum(dot4(row1, row2));
                        if (ff <= -(float)MAX_EXP || ff >= (float)MAX_EXP) continue;
                        const float g = (0.0f - s_exp[(int)((ff + (float)MAX_EXP) * (float)(EXP_TABLE_SIZE / MAX_EXP / 2))]) * alpha;
                        axpy4(work, g, row2);
                        float4 upd = row2;
                        axpy4(upd, g, row1);
                        add_row<ATOMIC>(rp, lane, make_float4(g * row1.x, g * row1.y, g * row1.z, g * row1.w), upd, on);
                    }
                }
                float4 upd1 = row1;
                upd1.x += work.x; upd1.y += work.y; upd1.z += work.z; upd1.w += work.w;
                add_row<ATOMIC>(row1p, lane, work, upd1, on);                  // syn0[ctx] += work
                ++pairs;
                t_cur = t_nxt;
            }
            if (carried >= 0) add_row<ATOMIC>(syn1neg + (int64_t)carried * dim, lane, pos_delta, pos_row, on);
            __syncwarp();
        }
    }
    if (lane == 0 && a.pairs_out && pairs) atomicAdd(a.pairs_out, pairs);
}

// ---- v3: one negative set per centre, shared by its context pairs ------------------------------
// (north-star subsystem 4: "staging of the shared negative set"; pWord2Vec-style sharing).
// The (1 positive + 5 negative) output rows of a centre are read ONCE, carried in registers across
// the centre's ~10 context pairs -- each pair runs fast_sentence_sg_neg's arithmetic in order
// against the carried rows -- and written back as one reduction per row. Per pair only the input
// row syn0[context] moves: 1,024 B + 6,144 B / pairs-per-centre instead of 7,168 B. Distribution
// of negatives per pair is unchanged (count^0.75); only their independence across one window is
// given up. Negative sets with a repeated row fall back to the uncarried sequential form.
#ifndef N2V_V3_MINB
#define N2V_V3_MINB 5
#endif
// FULL: dim == 128 exactly (every lane holds 4 floats of every row, no masking)
template <bool ATOMIC, bool FULL, bool SHARDED>
__global__ void __launch_bounds__(SGNS_BLOCK, N2V_V3_MINB)
sgns_train_kernel_v3(SgnsArgs a)
{
    constexpr int FN = 5;
    __shared__ int32_t s_idx[SGNS_BLOCK / 32][SGNS_SMEM_TOKENS];
    __shared__ uint16_t s_pos[SGNS_BLOCK / 32][SGNS_SMEM_TOKENS];
    __shared__ uint8_t s_rw[SGNS_BLOCK / 32][SGNS_SMEM_TOKENS];
    __shared__ float s_exp[EXP_TABLE_SIZE];
    // the carried rows as first read (one reduction of out - orig per row at the end of a centre):
    // parked in shared memory, each lane touches only its own 16 bytes -- keeps 24 registers free
    __shared__ float4 s_orig[SGNS_BLOCK / 32][FN + 1][32];
    __shared__ float *s_parts[2][8];
    for (int i = threadIdx.x; i < EXP_TABLE_SIZE; i += blockDim.x) s_exp[i] = exp_table_entry(i);
    if (SHARDED && threadIdx.x < 16) s_parts[threadIdx.x >> 3][threadIdx.x & 7] = (threadIdx.x < 8 ? a.parts0 : a.parts1)[threadIdx.x & 7];
    __syncthreads();

    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const WarpSentence ws{s_idx[wib], s_pos[wib], s_rw[wib]};
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = a.p.grid_warps;
    if (warp >= n_warps) return;
    const int32_t dim = FULL ? 128 : a.p.dim, window = a.p.window;
    const uint32_t k0 = (uint32_t)a.p.seed, k1 = (uint32_t)(a.p.seed >> 32);
    const uint32_t ep8 = a.p.epoch << 8;
    const bool on = FULL || (lane * 4 < dim);
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    using Rows = typename std::conditional<SHARDED, RowsSharded, RowsFlat>::type;
    Rows rows;
    if constexpr (SHARDED) rows = RowsSharded{s_parts[0], s_parts[1], dim, a.parts_log2, (1 << a.parts_log2) - 1};
    else rows = RowsFlat{a.syn0, a.syn1neg, dim};
    unsigned long long pairs = 0, centres = 0;

    // lane n draws shared negative n of centre position i: Philox ctr (pos_i << 16 | 0xFFFF)
    auto draw_centre = [&](int32_t i, uint64_t gs) -> int32_t {
        int32_t t = -1;
        if (lane < FN) {
            const Philox4 r = philox4x32_10((uint32_t)gs, (uint32_t)(gs >> 32), ((uint32_t)ws.pos[i] << 16) | 0xFFFFu,
                                            ep8 | (uint32_t)(1 + (lane >> 2)), k0, k1);
            const uint32_t rr = (lane & 3) == 0 ? r.x : (lane & 3) == 1 ? r.y : (lane & 3) == 2 ? r.z : r.w;
            t = draw_negative(rr, a.cum_table, a.bucket_lo, a.p.V, a.p.bucket_bits);
        }
        return t;
    };
    auto sigmoid_g = [&](float f, float label, float alpha) -> float {
        return (label - s_exp[(int)((f + (float)MAX_EXP) * (float)(EXP_TABLE_SIZE / MAX_EXP / 2))]) * alpha;
    };

    for (int64_t s = warp; s < a.n_sent; s += n_warps) {
        const int64_t tb = a.sent_off ? a.sent_off[s] : s * (int64_t)a.stride;
        int64_t tl = a.sent_off ? a.sent_off[s + 1] - tb : (int64_t)a.stride;
        if (tl > a.p.max_sentence_len) tl = a.p.max_sentence_len;
        const uint64_t gs = (uint64_t)(a.sent_id_base + s);
        const float alpha = job_alpha(a.p, s);
        int64_t t_next = 0;
        int32_t n_kept = 0, c_lo = 0, c_hi = 0;
        bool first_chunk = true;
        while (next_chunk(a, ws, tb, tl, t_next, gs, ep8, k0, k1, lane, n_kept, c_lo, c_hi, first_chunk)) {
            // centres that have at least one context pair, in order; negatives drawn one centre ahead
            auto bounds = [&](int32_t i, int32_t &j0, int32_t &kend) -> bool {
                j0 = i - window + ws.rw[i]; if (j0 < 0) j0 = 0;
                kend = i + window + 1 - ws.rw[i]; if (kend > n_kept) kend = n_kept;
                return (kend - j0) > ((i >= j0 && i < kend) ? 1 : 0);
            };
            int32_t i = c_lo, j0 = 0, kend = 0;
            while (i < c_hi && !bounds(i, j0, kend)) ++i;
            int32_t t_cur = i < c_hi ? draw_centre(i, gs) : -1;
            while (i < c_hi) {
                const int32_t centre = ws.idx[i];
                ++centres;
                int32_t tg[FN];
                bool dup = false;
#pragma unroll
                for (int d = 0; d < FN; ++d) tg[d] = __shfl_sync(0xFFFFFFFFu, t_cur, d);
#pragma unroll
                for (int d1 = 0; d1 < FN; ++d1)
#pragma unroll
                    for (int d2 = d1 + 1; d2 < FN; ++d2) dup |= (tg[d1] == tg[d2]) && (tg[d1] != centre);
                // next centre with pairs + its negatives (off the critical path)
                int32_t ni = i + 1, nj0 = 0, nkend = 0;
                while (ni < c_hi && !bounds(ni, nj0, nkend)) ++ni;
                if (!dup) {
                    float4 out[FN + 1];
                    uint32_t skipmask = 0xC0u;                 // padding targets 6, 7
                    out[0] = on ? ldcg4(rows.r1(centre), lane) : zero4;
#pragma unroll
                    for (int d = 0; d < FN; ++d)
                        out[d + 1] = (on && tg[d] != centre) ? ldcg4(rows.r1(tg[d]), lane) : zero4;
#pragma unroll
                    for (int d = 0; d < FN; ++d) if (tg[d] == centre) skipmask |= 2u << d;   // skipped, not redrawn
                    // hot negatives (the most frequent words): not carried -- reduced and re-read pair by pair,
                    // because hundreds of warps would hold stale copies of such a row at the same time
                    uint32_t hotmask = 0u;
#pragma unroll
                    for (int d = 0; d < FN; ++d) if (tg[d] < a.p.hot_rows && !((skipmask >> (d + 1)) & 1u)) hotmask |= 2u << d;
#pragma unroll
                    for (int d = 0; d <= FN; ++d) s_orig[wib][d][lane] = out[d];
                    int32_t j = (j0 == i) ? j0 + 1 : j0;
                    float4 row1 = on ? ldcg4(rows.r0(ws.idx[j]), lane) : zero4;
                    const int32_t t_nxt = ni < c_hi ? draw_centre(ni, gs) : -1;
                    if (ni < c_hi) {                       // next centre's output rows: L2 warm-up
                        prefetch_row_l2(rows.r1(ws.idx[ni]), lane);
#pragma unroll
                        for (int d = 0; d < FN; ++d) {
                            const int32_t tn = __shfl_sync(0xFFFFFFFFu, t_nxt, d);
                            if (on) prefetch_row_l2(rows.r1(tn), lane);
                        }
                    }
                    // the centre's pairs. Two instantiations: centres whose set holds a hot row (rare on large
                    // vocabularies) pay for the per-pair reductions / re-reads, the others run the lean loop
                    auto pair_loop = [&](auto hot_tag) {
                    constexpr bool HOT = decltype(hot_tag)::value;
                    while (j < kend) {
                        const int32_t ctx = ws.idx[j];
                        int32_t jn = j + 1; if (jn == i) ++jn;
                        // next input row always in flight (clamped past the window's end); it is
                        // stale only if it is the very row this pair is about to update
                        const int32_t ctx_n = ws.idx[jn < kend ? jn : j];
                        const bool stale = ctx_n == ctx;
                        const float4 row1n = on ? ldcg4(rows.r0(ctx_n), lane) : zero4;

                        // 6 dot products, reduced by a transposing butterfly: after the rounds on lane
                        // bits 4,3,2 each lane holds ONE of the (padded) 8 sums, bits 1,0 finish it:
                        // 9 shuffles instead of 30, and the sigmoid is evaluated once per target
                        // (in the 4 lanes that own it) instead of 6 times per lane.
                        float a0, a1, a2, a3;
                        {
                            const float p0 = dot4(row1, out[0]), p1 = dot4(row1, out[1]), p2 = dot4(row1, out[2]),
                                        p3 = dot4(row1, out[3]), p4 = dot4(row1, out[4]), p5 = dot4(row1, out[5]);
                            const bool h = lane & 16;
                            a0 = (h ? p4 : p0) + __shfl_xor_sync(0xFFFFFFFFu, h ? p0 : p4, 16);
                            a1 = (h ? p5 : p1) + __shfl_xor_sync(0xFFFFFFFFu, h ? p1 : p5, 16);
                            a2 = (h ? 0.f : p2) + __shfl_xor_sync(0xFFFFFFFFu, h ? p2 : 0.f, 16);
                            a3 = (h ? 0.f : p3) + __shfl_xor_sync(0xFFFFFFFFu, h ? p3 : 0.f, 16);
                        }
                        float fv;
                        {
                            const bool h8 = lane & 8, h4 = lane & 4;
                            const float b0 = (h8 ? a2 : a0) + __shfl_xor_sync(0xFFFFFFFFu, h8 ? a0 : a2, 8);
                            const float b1 = (h8 ? a3 : a1) + __shfl_xor_sync(0xFFFFFFFFu, h8 ? a1 : a3, 8);
                            fv = (h4 ? b1 : b0) + __shfl_xor_sync(0xFFFFFFFFu, h4 ? b0 : b1, 4);
                            fv += __shfl_xor_sync(0xFFFFFFFFu, fv, 2);
                            fv += __shfl_xor_sync(0xFFFFFFFFu, fv, 1);
                        }
                        // lane owns target v = lane >> 2 (0 = centre, 1..5 = negatives, 6,7 = padding)
                        float gv = 0.0f;
                        if (!((skipmask >> (lane >> 2)) & 1u) && fv > -(float)MAX_EXP && fv < (float)MAX_EXP)
                            gv = sigmoid_g(fv, lane < 4 ? 1.0f : 0.0f, alpha);
                        float4 work = zero4;
#pragma unroll
                        for (int d = 0; d <= FN; ++d) {       // g == 0: target skipped or |f| >= 6 (no-op)
                            const float g = __shfl_sync(0xFFFFFFFFu, gv, d * 4);
                            axpy4(work, g, out[d]);           // work += g * syn1neg[t]
                            if (HOT && d > 0 && ((hotmask >> d) & 1u)) {   // hot row: update now, take the row as it is now
                                float *const rp = rows.r1(tg[d - 1]);
                                axpy4(out[d], g, row1);
                                add_row<ATOMIC>(rp, lane, make_float4(g * row1.x, g * row1.y, g * row1.z, g * row1.w), out[d], on);
                                if (ATOMIC) out[d] = on ? ldcg4(rp, lane) : zero4;
                            } else {
                                axpy4(out[d], g, row1);       // syn1neg[t] += g * row1 (carried)
                            }
                        }
                        float4 upd1 = row1;
                        upd1.x += work.x; upd1.y += work.y; upd1.z += work.z; upd1.w += work.w;
                        add_row<ATOMIC>(rows.r0(ctx), lane, work, upd1, on);
                        j = jn;
                        row1 = row1n;
                        if (stale && j < kend) row1 = on ? ldcg4(rows.r0(ctx), lane) : zero4;   // re-read after the update
                    }
                    };
                    if (hotmask) pair_loop(std::true_type{}); else pair_loop(std::false_type{});
                    pairs += (unsigned long long)(kend - j0 - ((i >= j0 && i < kend) ? 1 : 0));
                    // one reduction per carried row: what this centre's pairs added to it
#pragma unroll
                    for (int d = 0; d <= FN; ++d) {
                        if (((skipmask | hotmask) >> d) & 1u) continue;
                        const float4 og = s_orig[wib][d][lane];
                        const float4 dl = make_float4(out[d].x - og.x, out[d].y - og.y, out[d].z - og.z, out[d].w - og.w);
                        add_row<ATOMIC>(rows.r1(d == 0 ? centre : tg[d - 1]), lane, dl, out[d], on);
                    }
                    t_cur = t_nxt;
                } else {
                    // repeated row in the set: uncarried sequential form (every target re-read per pair)
                    const int32_t t_nxt = ni < c_hi ? draw_centre(ni, gs) : -1;
                    const bool act1[1] = {on};
                    for (int32_t j = j0; j < kend; ++j) {
                        if (j == i) continue;
                        train_pair<1, ATOMIC>(rows, dim, centre, ws.idx[j], t_cur, FN, alpha, act1, s_exp, lane);
                        ++pairs;
                    }
                    t_cur = t_nxt;
                }
                i = ni; j0 = nj0; kend = nkend;
            }
            __syncwarp();
        }
    }
    if (lane == 0 && a.pairs_out && pairs) { atomicAdd(a.pairs_out, pairs); atomicAdd(a.pairs_out + 1, centres); }
}

int launch_train_mma(const SgnsArgs &a, cudaStream_t stream);     // n2v_sgns_mma.cu

template <int NV>
static int launch_train(const SgnsArgs &a, cudaStream_t stream)
{
    const int warps_per_block = SGNS_BLOCK / 32;
    const int blocks = (a.p.grid_warps + warps_per_block - 1) / warps_per_block;
    if (a.p.atomic_updates) sgns_train_kernel<NV, true><<<blocks, SGNS_BLOCK, 0, stream>>>(a);
    else sgns_train_kernel<NV, false><<<blocks, SGNS_BLOCK, 0, stream>>>(a);
    N2V_LAUNCH_CHECK();
    return N2V_OK;
}

template <int MINB>
static int launch_train_v2(const SgnsArgs &a, cudaStream_t stream)
{
    const int warps_per_block = SGNS_BLOCK / 32;
    const int blocks = (a.p.grid_warps + warps_per_block - 1) / warps_per_block;
    if (a.p.atomic_updates) sgns_train_kernel_v2<true, MINB><<<blocks, SGNS_BLOCK, 0, stream>>>(a);
    else sgns_train_kernel_v2<false, MINB><<<blocks, SGNS_BLOCK, 0, stream>>>(a);
    N2V_LAUNCH_CHECK();
    return N2V_OK;
}

static inline size_t align_up(size_t x, size_t al = 256) { return (x + al - 1) / al * al; }

}  // namespace n2v

using namespace n2v;

extern "C" int n2v_vocab_count(const int32_t *tokens, int64_t n_tokens, int32_t n_ids,
                               unsigned long long *counts, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    N2V_REQUIRE(n_tokens >= 0 && n_ids >= 0, "negative size");
    if (n_tokens == 0) return N2V_OK;
    N2V_REQUIRE(tokens && counts, "NULL buffer");
    int sms = sm_count();
    if (sms <= 0) { set_error("no CUDA device"); return N2V_ECUDA; }
    int64_t blocks = (n_tokens + 255) / 256;
    if (blocks > (int64_t)sms * 16) blocks = (int64_t)sms * 16;
    vocab_count_kernel<<<(unsigned)blocks, 256, 0, stream>>>(tokens, n_tokens, n_ids, counts);
    N2V_LAUNCH_CHECK();
    return N2V_OK;
}

extern "C" size_t n2v_sgns_prepare_workspace_bytes(int32_t V)
{
    size_t tb = 0;
    cub::DeviceScan::InclusiveSum(nullptr, tb, (double *)nullptr, (double *)nullptr, V);
    return align_up(tb) + 4 * align_up(sizeof(double) * (size_t)(V > 0 ? V : 1));
}

extern "C" int n2v_sgns_prepare(const unsigned long long *counts, int32_t V, double sample,
                                uint32_t *keep_thr, uint32_t *cum_table, int32_t *bucket_lo,
                                int32_t bucket_bits, void *workspace, size_t workspace_bytes,
                                void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    N2V_REQUIRE(V > 0, "empty vocabulary");
    N2V_REQUIRE(counts && keep_thr && cum_table && bucket_lo && workspace, "NULL buffer");
    N2V_REQUIRE(bucket_bits >= 1 && bucket_bits <= 24, "bucket_bits out of range");
    if (n2v_sgns_prepare_workspace_bytes(V) > workspace_bytes) { set_error("sgns_prepare workspace too small"); return N2V_ENOMEM; }
    size_t tb = 0;
    cub::DeviceScan::InclusiveSum(nullptr, tb, (double *)nullptr, (double *)nullptr, V);
    char *p = (char *)workspace;
    void *cub_tmp = p; p += align_up(tb);
    const size_t vb = align_up(sizeof(double) * (size_t)V);
    double *powed = (double *)p; p += vb;
    double *countd = (double *)p; p += vb;
    double *cum_pow = (double *)p; p += vb;
    double *cum_cnt = (double *)p;
    const int T = 256;
    sgns_pow_kernel<<<(V + T - 1) / T, T, 0, stream>>>(counts, V, powed, countd);
    N2V_LAUNCH_CHECK();
    size_t t1 = tb;
    N2V_CHECK_CUDA(cub::DeviceScan::InclusiveSum(cub_tmp, t1, powed, cum_pow, V, stream));
    t1 = tb;
    N2V_CHECK_CUDA(cub::DeviceScan::InclusiveSum(cub_tmp, t1, countd, cum_cnt, V, stream));
    sgns_prepare_kernel<<<(V + T - 1) / T, T, 0, stream>>>(counts, V, sample, cum_pow, cum_cnt, keep_thr, cum_table);
    N2V_LAUNCH_CHECK();
    const int nb = (1 << bucket_bits) + 1;
    sgns_bucket_kernel<<<(nb + T - 1) / T, T, 0, stream>>>(cum_table, V, bucket_lo, bucket_bits);
    N2V_LAUNCH_CHECK();
    return N2V_OK;
}

extern "C" int n2v_sgns_init(float *syn0, float *syn1neg, int32_t V, int32_t dim, uint64_t seed,
                             void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    N2V_REQUIRE(V >= 0 && dim > 0, "bad size");
    if (V == 0) return N2V_OK;
    N2V_REQUIRE(syn0, "syn0 is NULL");
    const int64_t n = (int64_t)V * ((dim + 3) >> 2);
    sgns_init_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(syn0, syn1neg, V, dim, (uint32_t)seed, (uint32_t)(seed >> 32), 0, 1);
    N2V_LAUNCH_CHECK();
    return N2V_OK;
}

extern "C" int n2v_sgns_init_part(float *syn0_part, float *syn1neg_part, int32_t V, int32_t dim, uint64_t seed,
                                  int32_t part, int32_t n_parts, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    N2V_REQUIRE(V >= 0 && dim > 0 && n_parts >= 1 && part >= 0 && part < n_parts, "bad size");
    const int32_t n_local = (V - part + n_parts - 1) / n_parts;
    if (n_local <= 0) return N2V_OK;
    N2V_REQUIRE(syn0_part, "syn0_part is NULL");
    const int64_t n = (int64_t)n_local * ((dim + 3) >> 2);
    sgns_init_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(syn0_part, syn1neg_part, n_local, dim, (uint32_t)seed,
                                                                      (uint32_t)(seed >> 32), part, n_parts);
    N2V_LAUNCH_CHECK();
    return N2V_OK;
}

static int sgns_train_impl(const int32_t *tokens, const int64_t *sent_off, int64_t n_sent,
                           int32_t stride, int64_t sent_id_base, const int32_t *vocab_of_id,
                           const uint32_t *keep_thr, const uint32_t *cum_table,
                           const int32_t *bucket_lo, const n2v_sgns_params_t *params,
                           float *syn0, float *syn1neg, float *const *parts0, float *const *parts1, int n_parts,
                           unsigned long long *pairs_out, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    N2V_REQUIRE(params, "params is NULL");
    N2V_REQUIRE(n_sent >= 0, "negative n_sent");
    if (n_sent == 0) return N2V_OK;
    const n2v_sgns_params_t &p = *params;
    N2V_REQUIRE(tokens && cum_table && bucket_lo && (n_parts > 0 || (syn0 && syn1neg)), "NULL buffer");
    N2V_REQUIRE(sent_off || stride > 0, "need sent_off or a positive stride");
    N2V_REQUIRE(p.V > 0 && p.dim > 0 && p.dim % 4 == 0 && p.dim <= 1024, "dim must be a multiple of 4, <= 1024");
    N2V_REQUIRE(p.window >= 1 && p.window <= SGNS_MAX_WINDOW, "window out of range (1..96)");
    N2V_REQUIRE(p.negative >= 0 && p.negative <= SGNS_MAX_NEG, "negative out of range");
    N2V_REQUIRE(p.max_sentence_len >= 1 && p.max_sentence_len <= 65535, "max_sentence_len out of range");
    N2V_REQUIRE(p.grid_warps >= 1 && p.total_examples >= 1 && p.sent_per_job >= 1, "bad schedule");
    N2V_REQUIRE(p.bucket_bits >= 1 && p.bucket_bits <= 24, "bucket_bits out of range");
    if (sm_count() <= 0) { set_error("no CUDA device"); return N2V_ECUDA; }
    SgnsArgs a{tokens, sent_off, n_sent, stride, sent_id_base, vocab_of_id, keep_thr, cum_table,
               bucket_lo, p, syn0, syn1neg, pairs_out, {}, {}, -1};
    if (n_parts > 0) {
        N2V_REQUIRE(n_parts <= 8 && (n_parts & (n_parts - 1)) == 0 && parts0 && parts1, "n_parts must be 1, 2, 4 or 8");
        N2V_REQUIRE(p.negative_sharing && p.dim <= 128 && p.negative == 5, "sharded tables: shared-negative kernel only");
        for (int i = 0; i < n_parts; ++i) {
            N2V_REQUIRE(parts0[i] && parts1[i], "NULL table part");
            a.parts0[i] = parts0[i]; a.parts1[i] = parts1[i];
        }
        a.parts_log2 = 0;
        while ((1 << a.parts_log2) < n_parts) ++a.parts_log2;
    }
    const int nv = (p.dim + 127) / 128;
    if (p.negative_sharing && (p.tuning & 16) && a.parts_log2 < 0) return launch_train_mma(a, stream);   // experiment (n2v_sgns_mma.cu)
    if (p.negative_sharing) {
        N2V_REQUIRE(nv == 1 && p.negative == 5, "negative_sharing needs dim <= 128 and negative == 5");
        const int blocks = (p.grid_warps + SGNS_BLOCK / 32 - 1) / (SGNS_BLOCK / 32);
        if (a.parts_log2 >= 0) {
            N2V_REQUIRE(p.atomic_updates, "sharded tables need atomic updates");
            if (p.dim == 128) sgns_train_kernel_v3<true, true, true><<<blocks, SGNS_BLOCK, 0, stream>>>(a);
            else sgns_train_kernel_v3<true, false, true><<<blocks, SGNS_BLOCK, 0, stream>>>(a);
        } else if (p.dim == 128) {
            if (p.atomic_updates) sgns_train_kernel_v3<true, true, false><<<blocks, SGNS_BLOCK, 0, stream>>>(a);
            else sgns_train_kernel_v3<false, true, false><<<blocks, SGNS_BLOCK, 0, stream>>>(a);
        } else {
            if (p.atomic_updates) sgns_train_kernel_v3<true, false, false><<<blocks, SGNS_BLOCK, 0, stream>>>(a);
            else sgns_train_kernel_v3<false, false, false><<<blocks, SGNS_BLOCK, 0, stream>>>(a);
        }
        N2V_LAUNCH_CHECK();
        return N2V_OK;
    }
    if (nv == 1 && p.negative == 5 && !(p.tuning & 8)) {      // reference configuration: v2
        switch (p.tuning & 3) {
            case 1: return launch_train_v2<4>(a, stream);
            case 2: return launch_train_v2<8>(a, stream);
            default: return launch_train_v2<6>(a, stream);
        }
    }
    switch (nv) {
        case 1: return launch_train<1>(a, stream);
        case 2: return launch_train<2>(a, stream);
        case 3: return launch_train<3>(a, stream);
        case 4: return launch_train<4>(a, stream);
        default: return launch_train<8>(a, stream);
    }
}

extern "C" int n2v_sgns_train(const int32_t *tokens, const int64_t *sent_off, int64_t n_sent,
                              int32_t stride, int64_t sent_id_base, const int32_t *vocab_of_id,
                              const uint32_t *keep_thr, const uint32_t *cum_table,
                              const int32_t *bucket_lo, const n2v_sgns_params_t *params,
                              float *syn0, float *syn1neg, unsigned long long *pairs_out,
                              void *stream)
{
    return sgns_train_impl(tokens, sent_off, n_sent, stride, sent_id_base, vocab_of_id, keep_thr, cum_table, bucket_lo,
                           params, syn0, syn1neg, nullptr, nullptr, 0, pairs_out, stream);
}

extern "C" int n2v_sgns_train_sharded(const int32_t *tokens, const int64_t *sent_off, int64_t n_sent,
                                      int32_t stride, int64_t sent_id_base, const int32_t *vocab_of_id,
                                      const uint32_t *keep_thr, const uint32_t *cum_table,
                                      const int32_t *bucket_lo, const n2v_sgns_params_t *params,
                                      float *const *syn0_parts, float *const *syn1neg_parts, int32_t n_parts,
                                      unsigned long long *pairs_out, void *stream)
{
    return sgns_train_impl(tokens, sent_off, n_sent, stride, sent_id_base, vocab_of_id, keep_thr, cum_table, bucket_lo,
                           params, nullptr, nullptr, syn0_parts, syn1neg_parts, n_parts, pairs_out, stream);
}
