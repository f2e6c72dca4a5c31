// Block-partitioned skip-gram negative sampling: the multi-GPU form of learn_embeddings
// (src/main.py:82-90, gensim 3.2.0 word2vec_inner.pyx fast_sentence_sg_neg arithmetic per pair).
//
// The two tables are cut into n_parts row sets (vocabulary row i -> part i % n_parts, local row
// i / n_parts; the vocabulary is sorted by count, so every part sees the same frequency profile).
// A pool of walks is expanded into its (centre, context) pairs, and the pairs whose centre lies in
// part k are bucketed by the part of their context: bucket (k, i) touches ONLY syn1neg part k
// (centre + negatives) and syn0 part i (context rows). n_parts buckets with pairwise different k and
// i are therefore independent: GPU k owns syn1neg part k for good, trains bucket (k, (k + e) %
// n_parts) in sub-step e and passes the syn0 part it holds round the ring between sub-steps. No row
// is ever replicated, so nothing has to be averaged.
//
//   n2v_sgns_pairs_count / _fill   walks -> pair streams of one centre part (count, scan, fill:
//                                  deterministic order = bucket, sentence, centre, context)
//   n2v_sgns_train_block           one bucket against (syn0 part, syn1neg part)
//
// Negatives: one set of 5 per RUN of `run_pairs` consecutive pairs of the stream (drawn from the
// count^0.75 table and mapped to the same-rank word of part k), carried in registers for the run
// together with the current centre row; a row repeated inside the set is used once, a negative
// equal to the pair's centre is skipped for that pair (gensim's rule). Per pair only syn0[context]
// moves: 1,024 B per pair + 1,024 B per carried row (centre changes + 5 per run; counted in
// pairs_out[1]).
#include <cub/cub.cuh>

#include "n2v_common.cuh"
#include "n2v_sgns_stage.cuh"

namespace n2v {

constexpr int BLK_MAX_PARTS = 8;
constexpr int BLK_FN = 5;

struct PairsArgs {
    SgnsArgs a;
    int32_t part, lg, n_parts, neg_group;      // neg_group G > 0: bit 31 of the centre field marks the first pair
                                               // (in its stream) of each block of G token positions of a walk
    int32_t *counts;            // [n_parts][n_sent]           (count pass)
    const int64_t *offsets;     // [n_parts][n_sent] exclusive (fill pass)
    uint2 *pairs; int64_t capacity;
    unsigned long long *overflow;
};

// One warp per sentence: sub-sample + window shrink exactly as the sentence-major kernels
// (load_chunk), then every centre of this part emits its window, bucketed by the context's part.
template <bool FILL>
__global__ void __launch_bounds__(SGNS_BLOCK)
sgns_pairs_kernel(PairsArgs g)
{
    __shared__ int32_t s_idx[SGNS_BLOCK / 32][SGNS_SMEM_TOKENS];
    __shared__ uint16_t s_pos[SGNS_BLOCK / 32][SGNS_SMEM_TOKENS];
    __shared__ uint8_t s_rw[SGNS_BLOCK / 32][SGNS_SMEM_TOKENS];
    __shared__ long long s_cur[SGNS_BLOCK / 32][BLK_MAX_PARTS];
    __shared__ int32_t s_last[SGNS_BLOCK / 32][BLK_MAX_PARTS];
    const SgnsArgs &a = g.a;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const WarpSentence ws{s_idx[wib], s_pos[wib], s_rw[wib]};
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int32_t window = a.p.window, mask = g.n_parts - 1;
    const uint32_t k0 = (uint32_t)a.p.seed, k1 = (uint32_t)(a.p.seed >> 32);
    const uint32_t ep8 = a.p.epoch << 8;
    const uint32_t lt = (1u << lane) - 1u;

    for (int64_t s = warp; s < a.n_sent; s += n_warps) {
        const int64_t tb = a.sent_off ? a.sent_off[s] : s * (int64_t)a.stride;
        int64_t tl = a.sent_off ? a.sent_off[s + 1] - tb : (int64_t)a.stride;
        if (tl > a.p.max_sentence_len) tl = a.p.max_sentence_len;
        const uint64_t gs = (uint64_t)(a.sent_id_base + s);
        if (lane < BLK_MAX_PARTS)
            s_cur[wib][lane] = (FILL && lane < g.n_parts) ? (long long)g.offsets[(int64_t)lane * a.n_sent + s] : 0ll;
        if (lane < BLK_MAX_PARTS) s_last[wib][lane] = -1;
        __syncwarp();
        int64_t t_next = 0;
        int32_t n_kept = 0, c_lo = 0, c_hi = 0;
        bool first_chunk = true;
        while (next_chunk(a, ws, tb, tl, t_next, gs, ep8, k0, k1, lane, n_kept, c_lo, c_hi, first_chunk)) {
            // the centres of this part, in order: 32 candidates per ballot, then only the set bits
            for (int32_t i0 = c_lo; i0 < c_hi; i0 += 32) {
              uint32_t todo = __ballot_sync(0xFFFFFFFFu, i0 + lane < c_hi && (ws.idx[i0 + lane] & mask) == g.part);
              while (todo) {
                const int32_t i = i0 + __ffs(todo) - 1;
                todo &= todo - 1;
                const int32_t centre = ws.idx[i];
                const int32_t grp = g.neg_group > 0 ? (int32_t)ws.pos[i] / g.neg_group : -1;
                int32_t j0 = i - window + ws.rw[i]; if (j0 < 0) j0 = 0;
                int32_t kend = i + window + 1 - ws.rw[i]; if (kend > n_kept) kend = n_kept;
                for (int32_t jb = j0; jb < kend; jb += 32) {
                    const int32_t j = jb + lane;
                    const bool valid = j < kend && j != i;
                    const int32_t x = valid ? ws.idx[j] : 0;
                    const int32_t b = valid ? (x & mask) : (BLK_MAX_PARTS + lane);
                    const uint32_t peers = __match_any_sync(0xFFFFFFFFu, b);
                    const int rank = __popc(peers & lt);
                    long long base = 0; int32_t lastg = -1;
                    if (valid) { base = s_cur[wib][b]; lastg = s_last[wib][b]; }
                    __syncwarp();
                    if (valid && rank == 0) { s_cur[wib][b] = base + __popc(peers); s_last[wib][b] = grp; }
                    __syncwarp();
                    if (FILL && valid) {
                        const long long o = base + rank;
                        const uint32_t flag = (g.neg_group > 0 && rank == 0 && lastg != grp) ? 0x80000000u : 0u;
                        if (o < g.capacity) g.pairs[o] = make_uint2((uint32_t)(centre >> g.lg) | flag, (uint32_t)(x >> g.lg));
                        else if (rank == 0) atomicAdd(g.overflow, 1ull);
                    }
                }
              }
            }
            __syncwarp();
        }
        if (!FILL && lane < g.n_parts) g.counts[(int64_t)lane * a.n_sent + s] = (int32_t)s_cur[wib][lane];
        __syncwarp();
    }
}

struct BlockArgs {
    const uint2 *pairs; int64_t n_pairs;
    float *syn0_part, *syn1neg_part;
    const uint32_t *cum_table; const int32_t *bucket_lo;
    int32_t V, dim, bucket_bits, part, lg, run_pairs, grid_warps;
    float alpha; uint64_t seed; uint32_t epoch, tag; int32_t cut, blocked;
    unsigned long long *pairs_out;
};

#ifndef N2V_BLK_MINB
#define N2V_BLK_MINB 5
#endif
// FULL: dim == 128 exactly (every lane holds 4 floats of every row, no masking)
template <bool ATOMIC, bool FULL>
__global__ void __launch_bounds__(SGNS_BLOCK, N2V_BLK_MINB)
sgns_block_kernel(BlockArgs a)
{
    constexpr int FN = BLK_FN;
    __shared__ float s_exp[EXP_TABLE_SIZE];
    __shared__ float4 s_orig[SGNS_BLOCK / 32][FN + 1][32];
    for (int i = threadIdx.x; i < EXP_TABLE_SIZE; i += blockDim.x) s_exp[i] = exp_table_entry(i);
    __syncthreads();

    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = a.grid_warps;
    if (warp >= n_warps) return;
    const int32_t dim = FULL ? 128 : a.dim, K = a.run_pairs;
    const bool on = FULL || (lane * 4 < dim);
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const uint32_t k0 = (uint32_t)a.seed, k1 = (uint32_t)(a.seed >> 32);
    const float alpha = a.alpha;
    float *const syn0 = a.syn0_part, *const syn1 = a.syn1neg_part;
    const int64_t n_runs = (a.n_pairs + K - 1) / K;
    unsigned long long pairs = 0, carried = 0;      // carried = output rows read into registers
    auto r0 = [&](int32_t l) -> float * { return syn0 + (int64_t)l * dim; };
    auto r1 = [&](int32_t l) -> float * { return syn1 + (int64_t)l * dim; };
    auto sigmoid_g = [&](float f, float label) -> float {
        return (label - s_exp[(int)((f + (float)MAX_EXP) * (float)(EXP_TABLE_SIZE / MAX_EXP / 2))]) * alpha;
    };

    // lane n (< 5) draws negative n of (run r, centre group `sub` of the run): sub = 0 is the run's own
    // set; with a.cut every centre change inside the run draws a fresh set (sub = index of the pair at
    // which the centre changes), so negatives are shared only by pairs of ONE centre (the law of the
    // sentence-major shared-negative kernel)
    auto draw_sub = [&](int64_t r, int32_t sub) -> int32_t {
        int32_t t = -1;
        if (lane < FN) {
            const Philox4 ph = philox4x32_10((uint32_t)r, (uint32_t)((uint64_t)r >> 32), a.tag,
                                             (a.epoch << 8) | ((uint32_t)sub << 2) | (uint32_t)(1 + (lane >> 2)), k0, k1);
            const uint32_t rr = (lane & 3) == 0 ? ph.x : (lane & 3) == 1 ? ph.y : (lane & 3) == 2 ? ph.z : ph.w;
            t = draw_negative(rr, a.cum_table, a.bucket_lo, a.V, a.bucket_bits) >> a.lg;
            if ((((int64_t)t << a.lg) | a.part) >= a.V) --t;
        }
        return t;
    };
    // runs of a warp: strided (run = warp, warp + n_warps, ...) or, a.blocked, one contiguous range per warp
    // (concurrent warps then work on far-apart walks, as the sentence-major kernels do)
    const int64_t per_warp = (n_runs + n_warps - 1) / n_warps;
    const int64_t run_lo = a.blocked ? warp * per_warp : warp;
    const int64_t run_hi = a.blocked ? (run_lo + per_warp < n_runs ? run_lo + per_warp : n_runs) : n_runs;
    const int64_t run_step = a.blocked ? 1 : n_warps;
    uint2 mine_next = (run_lo * K + lane < a.n_pairs && lane < K) ? __ldcs(a.pairs + run_lo * K + lane) : make_uint2(0u, 0u);
    for (int64_t run = run_lo; run < run_hi; run += run_step) {
        const int64_t p0 = run * K;
        const int32_t cnt = (int32_t)((a.n_pairs - p0) < K ? (a.n_pairs - p0) : K);
        uint2 mine = mine_next;
        {                                                        // the next run's pairs: in flight during this run
            const int64_t pn = (run + run_step) * K;
            mine_next = (pn + lane < a.n_pairs && lane < K) ? __ldcs(a.pairs + pn + lane) : make_uint2(0u, 0u);
        }
        const uint32_t flags = __ballot_sync(0xFFFFFFFFu, (mine.x >> 31) != 0u);   // pairs that open a negative group
        mine.x &= 0x7FFFFFFFu;
        int32_t tg[FN];
        uint32_t base_skip = 0xC0u;                            // padding targets 6, 7
        float4 out[FN + 1];
        out[0] = zero4;
        auto flush_negs = [&]() {
#pragma unroll
            for (int d = 1; d <= FN; ++d) {
                if ((base_skip >> d) & 1u) continue;
                const float4 og = s_orig[wib][d][lane];
                add_row<ATOMIC>(r1(tg[d - 1]), lane,
                                make_float4(out[d].x - og.x, out[d].y - og.y, out[d].z - og.z, out[d].w - og.w), out[d], on);
            }
        };
        auto load_negs = [&](int32_t sub) {
            const int32_t t_run = draw_sub(run, sub);
            base_skip = 0xC0u;
#pragma unroll
            for (int d = 0; d < FN; ++d) tg[d] = __shfl_sync(0xFFFFFFFFu, t_run, d);
#pragma unroll
            for (int d1 = 0; d1 < FN; ++d1)
#pragma unroll
                for (int d2 = d1 + 1; d2 < FN; ++d2) if (tg[d1] == tg[d2]) base_skip |= 2u << d2;   // repeated row: once
#pragma unroll
            for (int d = 0; d < FN; ++d)
                out[d + 1] = (on && !((base_skip >> (d + 1)) & 1u)) ? ldcg4(r1(tg[d]), lane) : zero4;
#pragma unroll
            for (int d = 1; d <= FN; ++d) s_orig[wib][d][lane] = out[d];
            carried += (unsigned long long)(FN - __popc(base_skip & 0x3Eu));
        };
        load_negs(0);
        int32_t cur_c = -1, ahead_c = -1;
        float4 ahead = zero4;                                   // the next centre's row, read one pair early
        uint32_t skipmask = base_skip;
        int32_t ctx = (int32_t)__shfl_sync(0xFFFFFFFFu, mine.y, 0);
        float4 row1 = on ? ldcg4(r0(ctx), lane) : zero4;
        for (int32_t q = 0; q < cnt; ++q) {
            const int32_t c = (int32_t)__shfl_sync(0xFFFFFFFFu, mine.x, q);
            const int32_t qn = q + 1 < cnt ? q + 1 : q;
            const int32_t ctx_n = (int32_t)__shfl_sync(0xFFFFFFFFu, mine.y, qn);
            const int32_t c_n = (int32_t)__shfl_sync(0xFFFFFFFFu, mine.x, qn);
            if (a.cut == 2 && q > 0 && ((flags >> q) & 1u)) {    // a new negative group starts at this pair
                flush_negs(); load_negs(q);
                skipmask = base_skip;
#pragma unroll
                for (int d = 0; d < FN; ++d) if (tg[d] == cur_c) skipmask |= 2u << d;
            }
            if (c != cur_c) {                                   // centre row: write back, take the next
                if (cur_c >= 0) {
                    const float4 og = s_orig[wib][0][lane];
                    add_row<ATOMIC>(r1(cur_c), lane, make_float4(out[0].x - og.x, out[0].y - og.y, out[0].z - og.z, out[0].w - og.w), out[0], on);
                    if (a.cut == 1) { flush_negs(); load_negs(q); ahead_c = -1; }
                }
                out[0] = ahead_c == c ? ahead : (on ? ldcg4(r1(c), lane) : zero4);
                s_orig[wib][0][lane] = out[0];
                cur_c = c;
                ++carried;
                skipmask = base_skip;
#pragma unroll
                for (int d = 0; d < FN; ++d) if (tg[d] == c) skipmask |= 2u << d;   // skipped, not redrawn
            }
            // the next centre row and the next input row are read one pair early (no write of this warp
            // can hit the former before use; the latter is stale only if it is the row this pair updates)
            const bool stale = ctx_n == ctx;
            if (c_n != c && !a.cut) { ahead = on ? ldcg4(r1(c_n), lane) : zero4; ahead_c = c_n; }
            const float4 row1n = on ? ldcg4(r0(ctx_n), lane) : zero4;

            // 6 dot products by the transposing butterfly of the sentence-major kernel (n2v_sgns.cu)
            float a0, a1, a2, a3;
            {
                const float d0 = dot4(row1, out[0]), d1 = dot4(row1, out[1]), d2 = dot4(row1, out[2]),
                            d3 = dot4(row1, out[3]), d4 = dot4(row1, out[4]), d5 = dot4(row1, out[5]);
                const bool h = lane & 16;
                a0 = (h ? d4 : d0) + __shfl_xor_sync(0xFFFFFFFFu, h ? d0 : d4, 16);
                a1 = (h ? d5 : d1) + __shfl_xor_sync(0xFFFFFFFFu, h ? d1 : d5, 16);
                a2 = (h ? 0.f : d2) + __shfl_xor_sync(0xFFFFFFFFu, h ? d2 : 0.f, 16);
                a3 = (h ? 0.f : d3) + __shfl_xor_sync(0xFFFFFFFFu, h ? d3 : 0.f, 16);
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
            float gv = 0.0f;      // lane owns target lane >> 2 (0 = centre, 1..5 = negatives, 6,7 = padding)
            if (!((skipmask >> (lane >> 2)) & 1u) && fv > -(float)MAX_EXP && fv < (float)MAX_EXP)
                gv = sigmoid_g(fv, lane < 4 ? 1.0f : 0.0f);
            float4 work = zero4;
#pragma unroll
            for (int d = 0; d <= FN; ++d) {
                const float gd = __shfl_sync(0xFFFFFFFFu, gv, d * 4);
                axpy4(work, gd, out[d]);
                axpy4(out[d], gd, row1);
            }
            float4 upd1 = row1;
            upd1.x += work.x; upd1.y += work.y; upd1.z += work.z; upd1.w += work.w;
            add_row<ATOMIC>(r0(ctx), lane, work, upd1, on);
            row1 = row1n;
            if (stale && q + 1 < cnt) row1 = on ? ldcg4(r0(ctx), lane) : zero4;   // re-read after the update
            ctx = ctx_n;
        }
        // one reduction per carried row: what this run added to it
        if (cur_c >= 0) {
            const float4 og = s_orig[wib][0][lane];
            add_row<ATOMIC>(r1(cur_c), lane, make_float4(out[0].x - og.x, out[0].y - og.y, out[0].z - og.z, out[0].w - og.w), out[0], on);
        }
        flush_negs();
        pairs += (unsigned long long)cnt;
        __syncwarp();
    }
    if (lane == 0 && a.pairs_out && pairs) { atomicAdd(a.pairs_out, pairs); atomicAdd(a.pairs_out + 1, carried); }
}

// ---- the same kernel with the look-ahead rows staged through shared memory by cp.async ------------
// ptxas gives the register look-ahead load of sgns_block_kernel the scoreboard the very next FMUL
// waits on (scripts/sass_scoreboards.py), so its latency is exposed in full; cp.async has no register
// destination and no scoreboard: the input row of pair q + 2 and, when the centre changes there, its
// centre row are copied global -> shared two pairs ahead (one commit group per pair, wait_group 1),
// every lane moving and later reading only its own 16 bytes (no warp synchronisation needed). A row
// fetched before one of this warp's own writes to it landed is replaced: an input row whose context
// equals that of one of the two previous pairs by the value that pair's update produced (kept in
// registers), a centre row that was flushed one pair back by a re-read.
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{
    const uint32_t sa = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

// HOTP (experiment, `tuning & 2`, off by default): the few most frequent input rows (local rows
// [0, HOT): the vocabulary is count-sorted) get their updates summed in a per-warp shared-memory slot
// and written back every HOT_FLUSH runs instead of one reduction per pair; the warp reads such a row
// fresh from memory and adds its own pending sum, so a one-warp run still equals the sequential law.
// Built to test whether the reductions on the hottest row are what makes the buckets of context part
// 0 slow (25 vs 17 ms, the critical path of the 8-GPU ring): they are not -- the buckets stay slow
// (so the reads / re-reads of that row remain suspect) and C2's AUC moves by up to +0.02
// (profiles/r01_v_*).
template <bool ATOMIC, bool FULL, bool HOTP>
__global__ void __launch_bounds__(SGNS_BLOCK, N2V_BLK_MINB)
sgns_block_kernel_async(BlockArgs a)
{
    constexpr int FN = BLK_FN;
    constexpr int W = SGNS_BLOCK / 32;
    constexpr int HOT = 4, HOT_FLUSH = 8;
    __shared__ float4 s_hot[HOTP ? W : 1][HOT][32];
    __shared__ float s_exp[EXP_TABLE_SIZE];
    __shared__ float4 s_orig[W][FN][32];           // the run's negative rows as first read
    __shared__ float4 s_ctx[W][4][32];             // ring: input rows of pairs q .. q + 2
    __shared__ float4 s_cen[W][4][32];             // ring: centre rows (current one = its first-read copy)
    for (int i = threadIdx.x; i < EXP_TABLE_SIZE; i += blockDim.x) s_exp[i] = exp_table_entry(i);
    __syncthreads();

    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = a.grid_warps;
    if (warp >= n_warps) return;
    const int32_t dim = FULL ? 128 : a.dim, K = a.run_pairs;
    const bool on = FULL || (lane * 4 < dim);
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const uint32_t k0 = (uint32_t)a.seed, k1 = (uint32_t)(a.seed >> 32);
    const float alpha = a.alpha;
    float *const syn0 = a.syn0_part, *const syn1 = a.syn1neg_part;
    const int64_t n_runs = (a.n_pairs + K - 1) / K;
    unsigned long long pairs = 0, carried = 0;
    auto r0 = [&](int32_t l) -> float * { return syn0 + (int64_t)l * dim; };
    auto r1 = [&](int32_t l) -> float * { return syn1 + (int64_t)l * dim; };
    auto sigmoid_g = [&](float f, float label) -> float {
        return (label - s_exp[(int)((f + (float)MAX_EXP) * (float)(EXP_TABLE_SIZE / MAX_EXP / 2))]) * alpha;
    };
    auto draw_run = [&](int64_t r) -> int32_t {
        int32_t t = -1;
        if (lane < FN) {
            const Philox4 ph = philox4x32_10((uint32_t)r, (uint32_t)((uint64_t)r >> 32), a.tag,
                                             (a.epoch << 8) | (uint32_t)(1 + (lane >> 2)), k0, k1);
            const uint32_t rr = (lane & 3) == 0 ? ph.x : (lane & 3) == 1 ? ph.y : (lane & 3) == 2 ? ph.z : ph.w;
            t = draw_negative(rr, a.cum_table, a.bucket_lo, a.V, a.bucket_bits) >> a.lg;
            if ((((int64_t)t << a.lg) | a.part) >= a.V) --t;
        }
        return t;
    };
    for (int d = 0; d < 4; ++d) { s_ctx[wib][d][lane] = zero4; s_cen[wib][d][lane] = zero4; }   // lanes beyond dim read zeros
    uint32_t hot_mask = 0;
    int32_t hot_runs = 0;
    if (HOTP) for (int h = 0; h < HOT; ++h) s_hot[wib][h][lane] = zero4;
    auto flush_hot = [&]() {
        for (int h = 0; h < HOT; ++h)
            if ((hot_mask >> h) & 1u) {
                const float4 sum = s_hot[wib][h][lane];
                add_row<ATOMIC>(r0(h), lane, sum, sum, on);
                s_hot[wib][h][lane] = zero4;
            }
        hot_mask = 0;
    };

    uint2 mine_next = (warp * K + lane < a.n_pairs && lane < K) ? __ldcs(a.pairs + warp * K + lane) : make_uint2(0u, 0u);
    for (int64_t run = warp; run < n_runs; run += n_warps) {
        const int64_t p0 = run * K;
        const int32_t cnt = (int32_t)((a.n_pairs - p0) < K ? (a.n_pairs - p0) : K);
        const uint2 mine = mine_next;
        {
            const int64_t pn = (run + n_warps) * K;
            mine_next = (pn + lane < a.n_pairs && lane < K) ? __ldcs(a.pairs + pn + lane) : make_uint2(0u, 0u);
        }
        // fetch(q): input row of pair q and, if the centre changes at q, its centre row; one group per pair
        uint32_t cen_issued = 0, cen_used = 0;
        auto fetch = [&](int32_t q) {
            if (q < cnt) {
                const int32_t x = (int32_t)__shfl_sync(0xFFFFFFFFu, mine.y, q);
                const int32_t c = (int32_t)__shfl_sync(0xFFFFFFFFu, mine.x, q);
                const int32_t cp = (int32_t)__shfl_sync(0xFFFFFFFFu, mine.x, q > 0 ? q - 1 : 0);
                if (on) cp_async16(&s_ctx[wib][q & 3][lane], reinterpret_cast<const float4 *>(r0(x)) + lane);
                if (q == 0 || c != cp) {
                    if (on) cp_async16(&s_cen[wib][cen_issued & 3][lane], reinterpret_cast<const float4 *>(r1(c)) + lane);
                    ++cen_issued;
                }
            }
            cp_async_commit();
        };
        fetch(0);
        fetch(1);
        const int32_t t_run = draw_run(run);
        int32_t tg[FN];
        uint32_t base_skip = 0xC0u;
#pragma unroll
        for (int d = 0; d < FN; ++d) tg[d] = __shfl_sync(0xFFFFFFFFu, t_run, d);
#pragma unroll
        for (int d1 = 0; d1 < FN; ++d1)
#pragma unroll
            for (int d2 = d1 + 1; d2 < FN; ++d2) if (tg[d1] == tg[d2]) base_skip |= 2u << d2;
        float4 out[FN + 1];
        out[0] = zero4;
#pragma unroll
        for (int d = 0; d < FN; ++d)
            out[d + 1] = (on && !((base_skip >> (d + 1)) & 1u)) ? ldcg4(r1(tg[d]), lane) : zero4;
#pragma unroll
        for (int d = 0; d < FN; ++d) s_orig[wib][d][lane] = out[d + 1];
        carried += (unsigned long long)(FN - __popc(base_skip & 0x3Eu));
        int32_t cur_c = -1, prev_c = -1;             // prev_c: the centre before cur_c (flushed when cur_c was taken)
        int32_t ctx_m1 = -1, ctx_m2 = -1;
        float4 upd_m1 = zero4, upd_m2 = zero4;       // the input rows of the two previous pairs after their update
        uint32_t skipmask = base_skip;
        for (int32_t q = 0; q < cnt; ++q) {
            cp_async_wait1();                          // all groups but pair q + 1's have landed
            const int32_t c = (int32_t)__shfl_sync(0xFFFFFFFFu, mine.x, q);
            const int32_t ctx = (int32_t)__shfl_sync(0xFFFFFFFFu, mine.y, q);
            if (c != cur_c) {                          // centre row: write back, take the next
                if (cur_c >= 0) {
                    const float4 og = s_cen[wib][(cen_used - 1) & 3][lane];
                    add_row<ATOMIC>(r1(cur_c), lane, make_float4(out[0].x - og.x, out[0].y - og.y, out[0].z - og.z, out[0].w - og.w), out[0], on);
                }
                float4 *slot = &s_cen[wib][cen_used & 3][lane];
                if (c == prev_c) { if (on) *slot = ldcg4(r1(c), lane); }   // fetched before its own flush landed
                out[0] = *slot;
                ++cen_used;
                prev_c = cur_c;
                cur_c = c;
                ++carried;
                skipmask = base_skip;
#pragma unroll
                for (int d = 0; d < FN; ++d) if (tg[d] == c) skipmask |= 2u << d;
            }
            fetch(q + 2);                              // after this pair's centre flush, before its input-row update
            float4 row1 = s_ctx[wib][q & 3][lane];
            const bool hot = HOTP && ctx < HOT;
            if (hot) {                                  // fresh + what this warp still holds back for that row
                row1 = on ? ldcg4(r0(ctx), lane) : zero4;
                const float4 pend = s_hot[wib][ctx][lane];
                row1.x += pend.x; row1.y += pend.y; row1.z += pend.z; row1.w += pend.w;
            } else if (ctx == ctx_m1) row1 = upd_m1;    // staged before this warp's own update of the row: take the
            else if (ctx == ctx_m2) row1 = upd_m2;      // value that update produced (what a re-read would return,
            ctx_m2 = ctx_m1; ctx_m1 = ctx;              // without a load queued behind the reductions on a hub row)

            float a0, a1, a2, a3;
            {
                const float d0 = dot4(row1, out[0]), d1 = dot4(row1, out[1]), d2 = dot4(row1, out[2]),
                            d3 = dot4(row1, out[3]), d4 = dot4(row1, out[4]), d5 = dot4(row1, out[5]);
                const bool h = lane & 16;
                a0 = (h ? d4 : d0) + __shfl_xor_sync(0xFFFFFFFFu, h ? d0 : d4, 16);
                a1 = (h ? d5 : d1) + __shfl_xor_sync(0xFFFFFFFFu, h ? d1 : d5, 16);
                a2 = (h ? 0.f : d2) + __shfl_xor_sync(0xFFFFFFFFu, h ? d2 : 0.f, 16);
                a3 = (h ? 0.f : d3) + __shfl_xor_sync(0xFFFFFFFFu, h ? d3 : 0.f, 16);
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
            float gv = 0.0f;
            if (!((skipmask >> (lane >> 2)) & 1u) && fv > -(float)MAX_EXP && fv < (float)MAX_EXP)
                gv = sigmoid_g(fv, lane < 4 ? 1.0f : 0.0f);
            float4 work = zero4;
#pragma unroll
            for (int d = 0; d <= FN; ++d) {
                const float gd = __shfl_sync(0xFFFFFFFFu, gv, d * 4);
                axpy4(work, gd, out[d]);
                axpy4(out[d], gd, row1);
            }
            float4 upd1 = row1;
            upd1.x += work.x; upd1.y += work.y; upd1.z += work.z; upd1.w += work.w;
            if (hot) {
                float4 pend = s_hot[wib][ctx][lane];
                pend.x += work.x; pend.y += work.y; pend.z += work.z; pend.w += work.w;
                s_hot[wib][ctx][lane] = pend;
                hot_mask |= 1u << ctx;
            } else {
                add_row<ATOMIC>(r0(ctx), lane, work, upd1, on);
            }
            upd_m2 = upd_m1; upd_m1 = upd1;
        }
        if (HOTP && ++hot_runs >= HOT_FLUSH) { flush_hot(); hot_runs = 0; }
        if (cur_c >= 0) {
            const float4 og = s_cen[wib][(cen_used - 1) & 3][lane];
            add_row<ATOMIC>(r1(cur_c), lane, make_float4(out[0].x - og.x, out[0].y - og.y, out[0].z - og.z, out[0].w - og.w), out[0], on);
        }
#pragma unroll
        for (int d = 1; d <= FN; ++d) {
            if ((base_skip >> d) & 1u) continue;
            const float4 og = s_orig[wib][d - 1][lane];
            add_row<ATOMIC>(r1(tg[d - 1]), lane,
                            make_float4(out[d].x - og.x, out[d].y - og.y, out[d].z - og.z, out[d].w - og.w), out[d], on);
        }
        pairs += (unsigned long long)cnt;
        __syncwarp();
    }
    if (HOTP) flush_hot();
    if (lane == 0 && a.pairs_out && pairs) { atomicAdd(a.pairs_out, pairs); atomicAdd(a.pairs_out + 1, carried); }
}

static inline size_t blk_align(size_t x, size_t al = 256) { return (x + al - 1) / al * al; }

static int log2_parts(int32_t n_parts)
{
    return n_parts == 1 ? 0 : n_parts == 2 ? 1 : n_parts == 4 ? 2 : n_parts == 8 ? 3 : -1;
}

static size_t pairs_scan_bytes(int64_t n)
{
    size_t b = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, b, (const int32_t *)nullptr, (int64_t *)nullptr, n);
    return blk_align(b);
}

static int pairs_args(PairsArgs &g, const int32_t *tokens, const int64_t *sent_off, int64_t n_sent, int32_t stride,
                      int64_t sent_id_base, const int32_t *vocab_of_id, const uint32_t *keep_thr,
                      const n2v_sgns_params_t *params, int32_t part, int32_t n_parts)
{
    N2V_REQUIRE(params, "params is NULL");
    N2V_REQUIRE(tokens && n_sent >= 0, "bad corpus");
    N2V_REQUIRE(sent_off || stride > 0, "sent_off is NULL and stride <= 0");
    N2V_REQUIRE(params->window >= 1 && params->window <= SGNS_MAX_WINDOW, "window out of range (1..96)");
    N2V_REQUIRE(params->max_sentence_len >= 1 && params->max_sentence_len <= 65535, "max_sentence_len out of range");
    const int lg = log2_parts(n_parts);
    N2V_REQUIRE(lg >= 0 && part >= 0 && part < n_parts, "n_parts must be 1, 2, 4 or 8 and 0 <= part < n_parts");
    memset(&g, 0, sizeof(g));
    g.a.tokens = tokens; g.a.sent_off = sent_off; g.a.n_sent = n_sent; g.a.stride = stride;
    g.a.sent_id_base = sent_id_base; g.a.vocab_of_id = vocab_of_id; g.a.keep_thr = keep_thr; g.a.p = *params;
    g.part = part; g.lg = lg; g.n_parts = n_parts;
    g.neg_group = (params->tuning >> 8) & 0xFF;
    return N2V_OK;
}

}  // namespace n2v

using namespace n2v;

extern "C" size_t n2v_sgns_pairs_workspace_bytes(int64_t n_sent, int32_t n_parts)
{
    const int64_t n = (int64_t)n_parts * (n_sent > 0 ? n_sent : 0) + 1;
    return blk_align(sizeof(int32_t) * (size_t)n) + pairs_scan_bytes(n);
}

extern "C" int n2v_sgns_pairs_count(const int32_t *tokens, const int64_t *sent_off, int64_t n_sent, int32_t stride,
                                    int64_t sent_id_base, const int32_t *vocab_of_id, const uint32_t *keep_thr,
                                    const n2v_sgns_params_t *params, int32_t part, int32_t n_parts,
                                    int64_t *offsets, void *workspace, size_t workspace_bytes, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    PairsArgs g;
    int rc = pairs_args(g, tokens, sent_off, n_sent, stride, sent_id_base, vocab_of_id, keep_thr, params, part, n_parts);
    if (rc != N2V_OK) return rc;
    N2V_REQUIRE(offsets && workspace, "offsets / workspace is NULL");
    const int64_t n = (int64_t)n_parts * n_sent + 1;
    N2V_REQUIRE(workspace_bytes >= n2v_sgns_pairs_workspace_bytes(n_sent, n_parts), "pairs workspace too small");
    g.counts = (int32_t *)workspace;
    void *tmp = (char *)workspace + blk_align(sizeof(int32_t) * (size_t)n);
    size_t tmp_bytes = pairs_scan_bytes(n);
    N2V_CHECK_CUDA(cudaMemsetAsync(g.counts + (n - 1), 0, sizeof(int32_t), stream));
    if (n_sent > 0) {
        int sms = sm_count();
        if (sms <= 0) { set_error("no CUDA device"); return N2V_ECUDA; }
        int64_t blocks = (n_sent + SGNS_BLOCK / 32 - 1) / (SGNS_BLOCK / 32);
        if (blocks > (int64_t)sms * 16) blocks = (int64_t)sms * 16;
        sgns_pairs_kernel<false><<<(unsigned)blocks, SGNS_BLOCK, 0, stream>>>(g);
        N2V_LAUNCH_CHECK();
    }
    N2V_CHECK_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, (const int32_t *)g.counts, offsets, n, stream));
    return N2V_OK;
}

extern "C" int n2v_sgns_pairs_fill(const int32_t *tokens, const int64_t *sent_off, int64_t n_sent, int32_t stride,
                                   int64_t sent_id_base, const int32_t *vocab_of_id, const uint32_t *keep_thr,
                                   const n2v_sgns_params_t *params, int32_t part, int32_t n_parts,
                                   const int64_t *offsets, int32_t *pairs, int64_t capacity_pairs,
                                   unsigned long long *overflow, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    PairsArgs g;
    int rc = pairs_args(g, tokens, sent_off, n_sent, stride, sent_id_base, vocab_of_id, keep_thr, params, part, n_parts);
    if (rc != N2V_OK) return rc;
    N2V_REQUIRE(offsets && pairs && overflow && capacity_pairs >= 0, "offsets / pairs / overflow is NULL");
    if (n_sent == 0) return N2V_OK;
    g.offsets = offsets; g.pairs = (uint2 *)pairs; g.capacity = capacity_pairs; g.overflow = overflow;
    int sms = sm_count();
    if (sms <= 0) { set_error("no CUDA device"); return N2V_ECUDA; }
    int64_t blocks = (n_sent + SGNS_BLOCK / 32 - 1) / (SGNS_BLOCK / 32);
    if (blocks > (int64_t)sms * 16) blocks = (int64_t)sms * 16;
    sgns_pairs_kernel<true><<<(unsigned)blocks, SGNS_BLOCK, 0, stream>>>(g);
    N2V_LAUNCH_CHECK();
    return N2V_OK;
}

extern "C" int n2v_sgns_train_block(const int32_t *pairs, int64_t n_pairs, const uint32_t *cum_table,
                                    const int32_t *bucket_lo, const n2v_sgns_params_t *params, float alpha,
                                    int32_t run_pairs, uint32_t tag, float *syn0_part, float *syn1neg_part,
                                    int32_t part, int32_t n_parts, unsigned long long *pairs_out, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    N2V_REQUIRE(params, "params is NULL");
    N2V_REQUIRE(n_pairs >= 0, "negative pair count");
    if (n_pairs == 0) return N2V_OK;
    N2V_REQUIRE(pairs && cum_table && bucket_lo && syn0_part && syn1neg_part, "NULL buffer");
    const int lg = log2_parts(n_parts);
    N2V_REQUIRE(lg >= 0 && part >= 0 && part < n_parts, "n_parts must be 1, 2, 4 or 8 and 0 <= part < n_parts");
    N2V_REQUIRE(params->V >= n_parts, "fewer vocabulary rows than parts");
    N2V_REQUIRE(params->dim >= 4 && params->dim <= 128 && params->dim % 4 == 0, "block kernel: dim must be a multiple of 4, <= 128");
    N2V_REQUIRE(params->negative == BLK_FN, "block kernel: negative must be 5");
    N2V_REQUIRE(run_pairs >= 1 && run_pairs <= 32, "run_pairs must be in [1, 32]");
    N2V_REQUIRE(params->bucket_bits >= 0 && params->bucket_bits <= 24, "bucket_bits out of range");
    N2V_REQUIRE(params->grid_warps >= 1, "grid_warps must be >= 1");
    BlockArgs a;
    memset(&a, 0, sizeof(a));
    a.pairs = (const uint2 *)pairs; a.n_pairs = n_pairs; a.syn0_part = syn0_part; a.syn1neg_part = syn1neg_part;
    a.cum_table = cum_table; a.bucket_lo = bucket_lo; a.V = params->V; a.dim = params->dim;
    a.bucket_bits = params->bucket_bits; a.part = part; a.lg = lg; a.run_pairs = run_pairs;
    a.alpha = alpha; a.seed = params->seed; a.epoch = params->epoch; a.tag = tag; a.pairs_out = pairs_out;
    const int64_t n_runs = (n_pairs + run_pairs - 1) / run_pairs;
    a.grid_warps = (int32_t)(n_runs < params->grid_warps ? n_runs : params->grid_warps);
    const int wpb = SGNS_BLOCK / 32;
    const int blocks = (a.grid_warps + wpb - 1) / wpb;
    const bool full = params->dim == 128;
    const bool reg_lookahead = (params->tuning & 1) != 0;     // 1: register look-ahead kernel (kept for comparison)
    const bool hot_private = (params->tuning & 2) != 0;       // 2: per-warp sums for the hottest input rows
    // 4 (register look-ahead kernel only): a fresh negative set at every centre change, or -- when the pair
    // streams carry group flags (bits 8-15 = G) -- at every flagged pair; 8: one contiguous range of runs per warp
    a.cut = (params->tuning & 4) ? (((params->tuning >> 8) & 0xFF) ? 2 : 1) : 0;
    a.blocked = (params->tuning & 8) ? 1 : 0;
    if (params->atomic_updates && !reg_lookahead && !a.cut && !a.blocked) {
        if (full && hot_private) sgns_block_kernel_async<true, true, true><<<blocks, SGNS_BLOCK, 0, stream>>>(a);
        else if (full) sgns_block_kernel_async<true, true, false><<<blocks, SGNS_BLOCK, 0, stream>>>(a);
        else if (hot_private) sgns_block_kernel_async<true, false, true><<<blocks, SGNS_BLOCK, 0, stream>>>(a);
        else sgns_block_kernel_async<true, false, false><<<blocks, SGNS_BLOCK, 0, stream>>>(a);
    } else if (params->atomic_updates) {
        if (full) sgns_block_kernel<true, true><<<blocks, SGNS_BLOCK, 0, stream>>>(a);
        else sgns_block_kernel<true, false><<<blocks, SGNS_BLOCK, 0, stream>>>(a);
    } else {
        if (full) sgns_block_kernel<false, true><<<blocks, SGNS_BLOCK, 0, stream>>>(a);
        else sgns_block_kernel<false, false><<<blocks, SGNS_BLOCK, 0, stream>>>(a);
    }
    N2V_LAUNCH_CHECK();
    return N2V_OK;
}
