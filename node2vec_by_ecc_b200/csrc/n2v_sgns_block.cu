// Block-partitioned skip-gram negative sampling: the multi-GPU form of learn_embeddings
// (src/main.py:82-90, gensim 3.2.0 word2vec_inner.pyx fast_sentence_sg_neg arithmetic per pair).
//
// The two tables are cut into n_parts row sets (vocabulary row i -> part i % n_parts, local row
// i / n_parts; the vocabulary is sorted by count, so every part sees the same frequency profile).
// A pool of walks is expanded into its (centre, context) pairs, and the pairs whose centre lies in
// part k are bucketed by the part of their context: bucket (k, b) touches ONLY syn1neg part k
// (centre + negatives) and syn0 part b (context rows). n_parts buckets with pairwise different k and
// b are therefore independent: GPU k owns syn1neg part k for good, trains bucket (k, (k + e) %
// n_parts) in sub-step e and passes the syn0 part it holds round the ring between sub-steps. No row
// is ever replicated, so nothing has to be averaged.
//
//   n2v_sgns_groups_count / _fill   walks -> the streams of one centre part (count, scan, fill:
//                                   deterministic order = stream, sentence, centre, context)
//   n2v_sgns_train_groups           one stream against (syn0 part, syn1neg part)
//
// Stream format (uint32 words). The pairs of one centre occurrence that fall into a stream form a
// GROUP: 8 header words {0x80000000 | centre local row, sentence index in the pool, token position |
// pair count << 16, the 5 negatives of the centre as local rows of its part} followed by one word per
// pair, the context's local row. Only the first header word has bit 31 set, so any word offset can be
// re-synchronised by scanning for it. The negatives are drawn ONCE per centre occurrence, by the
// expansion kernel, and repeated in the header of every stream the centre reaches: the training kernel
// (whose groups hold ~2 pairs at 8 parts) runs no Philox rounds and no table search of its own.
//
// Law. Exactly the sentence-major shared-negative kernel's (n2v_sgns.cu v3), whatever n_parts is: ONE
// negative set per centre occurrence, drawn from Philox (sentence id, position) -> count^0.75 table --
// the same draws that kernel makes -- and mapped to the word of the same local row in the centre's part
// (negatives must live where the centre lives); alpha follows the sentence's job exactly as there. The
// negatives of a centre are thus the same in every stream it appears in, and nothing about the law
// depends on the number of GPUs. (Round 1 shared one set per run of 32 consecutive pairs of a stream,
// i.e. across different centres, and gave consecutive runs to different warps: on C2 that moved the
// link-prediction AUC by +0.008 at 1 part and -0.006 at 4, profiles/r02_a_block_auc_bisect.txt.)
// neg_group = G > 1 lets G consecutive token positions of a walk share a set (fewer row operations per
// pair at large n_parts; a different estimator, measured below the band on C2 -- off by default).
//
// Work distribution: warp w owns one contiguous range of the stream (whole groups), i.e. consecutive
// walks, so concurrent warps work on far-apart walks as the sentence-major kernels do.
#include <cub/cub.cuh>
#include <type_traits>

#include "n2v_common.cuh"
#include "n2v_sgns_stage.cuh"

namespace n2v {

constexpr int BLK_MAX_PARTS = 8;
constexpr int BLK_FN = 5;
constexpr uint32_t GROUP_FLAG = 0x80000000u;
constexpr int GROUP_HDR = 8;             // header words of a group

struct GroupsArgs {
    SgnsArgs a;
    int32_t part, lg, n_parts;
    int32_t *counts;            // [n_parts][n_sent] words    (count pass)
    const int64_t *offsets;     // [n_parts][n_sent] exclusive (fill pass)
    uint32_t *words; int64_t capacity;
    unsigned long long *overflow;
    const uint32_t *cum_table; const int32_t *bucket_lo; int32_t neg_group;     // fill pass: the centres' negative sets
};

constexpr int BLK_STAGE_WORDS = 1024;    // per-warp staging of one batch of groups (fill pass)
#ifndef N2V_GROUPS_MINB
#define N2V_GROUPS_MINB 7                // the expansion is latency-bound: 7 blocks = 28 warps per SM (29 KB, <= 73 registers)
#endif

// One warp per sentence: sub-sample + window shrink exactly as the sentence-major kernels (load_chunk);
// then ONE LANE PER CENTRE of this part, a batch of up to 32 centres at a time: every lane walks its own
// window and counts its contexts per stream (8 x 8-bit fields), a warp scan of the per-stream group sizes
// (16-bit fields) places every group, and the 5 negatives of all centres of the batch are drawn side by
// side (Philox and table search spread over the lanes). The fill pass assembles the batch's groups in
// shared memory, stream by stream, and copies each stream's run out with coalesced stores. The order
// inside a stream is (sentence, centre position, context position), the same whatever the batching.
template <bool FILL>
__global__ void __launch_bounds__(SGNS_BLOCK, N2V_GROUPS_MINB)
sgns_groups_kernel(GroupsArgs g)
{
    constexpr int WPB = SGNS_BLOCK / 32;
    constexpr uint32_t FULLM = 0xFFFFFFFFu;
    static_assert(SGNS_SMEM_TOKENS <= 256 && BLK_STAGE_WORDS + GROUP_HDR < 65536, "s_list holds 8-bit positions, s_off 16-bit slots");
    __shared__ int32_t s_idx[WPB][SGNS_SMEM_TOKENS];
    __shared__ uint16_t s_pos[WPB][SGNS_SMEM_TOKENS];
    __shared__ uint8_t s_rw[WPB][SGNS_SMEM_TOKENS];
    __shared__ uint8_t s_list[WPB][SGNS_SMEM_TOKENS];                            // the chunk's centres of this part
    __shared__ uint32_t s_stage[FILL ? WPB : 1][FILL ? BLK_STAGE_WORDS : 1];
    __shared__ uint16_t s_off[FILL ? WPB : 1][BLK_MAX_PARTS][FILL ? 32 : 1];     // [stream][lane]: next context slot
    __shared__ uint32_t s_neg[FILL ? WPB : 1][FILL ? 32 * BLK_FN : 1];           // [centre of the batch][negative]
    const SgnsArgs &a = g.a;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const WarpSentence ws{s_idx[wib], s_pos[wib], s_rw[wib]};
    uint8_t *const list = s_list[wib];
    uint32_t *const stage = s_stage[FILL ? wib : 0];
    uint32_t *const neg = s_neg[FILL ? wib : 0];
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int32_t window = a.p.window, n_parts = g.n_parts, mask = n_parts - 1;
    const uint32_t k0 = (uint32_t)a.p.seed, k1 = (uint32_t)(a.p.seed >> 32);
    const uint32_t ep8 = a.p.epoch << 8;
    const uint32_t lt = (1u << lane) - 1u;
    int32_t batch = 32;
    if (FILL) { const int32_t most = BLK_STAGE_WORDS / (2 * window + GROUP_HDR * n_parts); if (most < batch) batch = most; }
    bool overflowed = false;

    for (int64_t s = warp; s < a.n_sent; s += n_warps) {
        const int64_t tb = a.sent_off ? a.sent_off[s] : s * (int64_t)a.stride;
        int64_t tl = a.sent_off ? a.sent_off[s + 1] - tb : (int64_t)a.stride;
        if (tl > a.p.max_sentence_len) tl = a.p.max_sentence_len;
        const uint64_t gs = (uint64_t)(a.sent_id_base + s);
        // lane b < n_parts keeps stream b's state: where this sentence's words start, words emitted so far
        long long base = 0; uint32_t cur = 0;
        if (FILL && lane < n_parts) base = (long long)g.offsets[(int64_t)lane * a.n_sent + s];
        int64_t t_next = 0;
        int32_t n_kept = 0, c_lo = 0, c_hi = 0;
        bool first_chunk = true;
        while (next_chunk(a, ws, tb, tl, t_next, gs, ep8, k0, k1, lane, n_kept, c_lo, c_hi, first_chunk)) {
            int32_t n_mine = 0;
            for (int32_t i0 = c_lo; i0 < c_hi; i0 += 32) {
                const bool mine = i0 + lane < c_hi && (ws.idx[i0 + lane] & mask) == g.part;
                const uint32_t m = __ballot_sync(FULLM, mine);
                if (mine) list[n_mine + __popc(m & lt)] = (uint8_t)(i0 + lane);
                n_mine += __popc(m);
            }
            __syncwarp();
            for (int32_t c0 = 0; c0 < n_mine; c0 += batch) {
                const int32_t nb = n_mine - c0 < batch ? n_mine - c0 : batch;
                const bool have = lane < nb;
                const int32_t i = have ? (int32_t)list[c0 + lane] : 0;
                int32_t j0 = 0, kend = 0;
                if (have) {
                    j0 = i - window + ws.rw[i]; if (j0 < 0) j0 = 0;
                    kend = i + window + 1 - ws.rw[i]; if (kend > n_kept) kend = n_kept;
                }
                // contexts per stream: 8-bit fields (a window holds <= 2 * 96 of them), streams 0-3 | 4-7
                uint32_t cl = 0, ch = 0;
                for (int32_t j = j0; j < kend; ++j) {
                    if (j == i) continue;
                    const int32_t b = ws.idx[j] & mask;
                    if (b < 4) cl += 1u << (8 * b); else ch += 1u << (8 * (b - 4));
                }
                // group sizes in words, 16-bit fields, two streams per register; inclusive scan over the batch
                // (32 groups of <= 200 words: no field overflows, and inc >= w field by field)
                uint32_t w[4], inc[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint32_t c = q < 2 ? cl >> (16 * q) : ch >> (16 * (q - 2));
                    const uint32_t c_even = c & 0xFFu, c_odd = (c >> 8) & 0xFFu;
                    w[q] = (c_even ? c_even + GROUP_HDR : 0u) | ((c_odd ? c_odd + GROUP_HDR : 0u) << 16);
                    inc[q] = w[q];
                }
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        if (2 * q < n_parts) {
                            const uint32_t t = __shfl_up_sync(FULLM, inc[q], d);
                            if (lane >= d) inc[q] += t;
                        }
                }
                // lane b: words the batch adds to stream b
                uint32_t tot = 0;
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (2 * q < n_parts) {
                        const uint32_t t = __shfl_sync(FULLM, inc[q], 31);
                        if ((lane >> 1) == q) tot = (lane & 1) ? t >> 16 : t & 0xFFFFu;
                    }
                if (lane >= n_parts) tot = 0;
                if (FILL) {
                    // negatives of the batch: Philox counters (gs lo, gs hi, position key << 16 | 0xFFFF, epoch << 8 |
                    // 1 + n / 4) as the sentence-major kernel's draw_centre, two blocks per centre (negatives 0-3, 4)
                    for (int32_t t = lane; t < 2 * nb; t += 32) {
                        const int32_t slot = t >> 1, which = t & 1;
                        const uint32_t pp = (uint32_t)ws.pos[list[c0 + slot]];
                        const uint32_t poskey = g.neg_group > 1 ? pp / (uint32_t)g.neg_group : pp;
                        const Philox4 r = philox4x32_10((uint32_t)gs, (uint32_t)(gs >> 32), (poskey << 16) | 0xFFFFu,
                                                        ep8 | (uint32_t)(1 + which), k0, k1);
                        uint32_t *o = neg + slot * BLK_FN;
                        if (which == 0) { o[0] = r.x; o[1] = r.y; o[2] = r.z; o[3] = r.w; } else o[4] = r.x;
                    }
                    __syncwarp();
                    // -> count^0.75 table -> the word of the same local row in this part
                    for (int32_t t = lane; t < BLK_FN * nb; t += 32) {
                        int32_t x = draw_negative(neg[t], g.cum_table, g.bucket_lo, a.p.V, a.p.bucket_bits) >> g.lg;
                        if ((((int64_t)x << g.lg) | g.part) >= a.p.V) --x;
                        neg[t] = (uint32_t)x;
                    }
                    // where stream b's run of this batch starts in the staging buffer
                    uint32_t run = tot;
#pragma unroll
                    for (int d = 1; d < BLK_MAX_PARTS; d <<= 1) {
                        const uint32_t t = __shfl_up_sync(FULLM, run, d);
                        if (lane >= d) run += t;
                    }
                    run -= tot;
                    __syncwarp();
                    // headers, and every lane's first context slot per stream
                    const uint32_t centre_row = have ? (uint32_t)(ws.idx[i] >> g.lg) : 0u;
                    const uint32_t pos_i = have ? (uint32_t)ws.pos[i] : 0u;
#pragma unroll
                    for (int b = 0; b < BLK_MAX_PARTS; ++b)
                        if (b < n_parts) {
                            const uint32_t so = __shfl_sync(FULLM, run, b);
                            const uint32_t cb = ((b < 4 ? cl : ch) >> (8 * (b & 3))) & 0xFFu;
                            const uint32_t ex = ((inc[b >> 1] - w[b >> 1]) >> (16 * (b & 1))) & 0xFFFFu;
                            if (have && cb) {
                                uint32_t *h = stage + so + ex;
                                h[0] = GROUP_FLAG | centre_row;
                                h[1] = (uint32_t)s;
                                h[2] = pos_i | (cb << 16);
#pragma unroll
                                for (int d = 0; d < BLK_FN; ++d) h[3 + d] = neg[lane * BLK_FN + d];
                                s_off[wib][b][lane] = (uint16_t)(so + ex + GROUP_HDR);
                            }
                        }
                    for (int32_t j = j0; j < kend; ++j) {
                        if (j == i) continue;
                        const int32_t x = ws.idx[j];
                        const uint32_t p = s_off[wib][x & mask][lane];
                        s_off[wib][x & mask][lane] = (uint16_t)(p + 1);
                        stage[p] = (uint32_t)(x >> g.lg);
                    }
                    __syncwarp();
                    // copy every stream's run out
                    for (int b = 0; b < n_parts; ++b) {
                        const uint32_t n_w = __shfl_sync(FULLM, tot, b), so = __shfl_sync(FULLM, run, b);
                        const long long go = __shfl_sync(FULLM, base + (long long)cur, b);
                        const long long room = g.capacity - go;
                        const uint32_t fit = room >= (long long)n_w ? n_w : room > 0 ? (uint32_t)room : 0u;
                        if (fit < n_w) overflowed = true;
                        uint32_t *const out = g.words + go;
                        const uint32_t *const in = stage + so;
                        for (uint32_t k = lane; k < fit; k += 32) out[k] = in[k];
                    }
                    __syncwarp();
                }
                cur += tot;
            }
            __syncwarp();
        }
        if (!FILL && lane < n_parts) g.counts[(int64_t)lane * a.n_sent + s] = (int32_t)cur;
        __syncwarp();
    }
    if (FILL && overflowed) atomicAdd(g.overflow, 1ull);
}

struct TrainGroupsArgs {
    const uint32_t *words;
    int64_t first, end, capacity;              // the stream = words [first, end), clamped to capacity
    const int64_t *dev_first, *dev_end;        // optional device copies of first / end (no host round trip)
    float *syn0_part, *syn1neg_part;
    n2v_sgns_params_t p;
    int32_t part, lg, neg_group;
    unsigned long long *pairs_out;
};

#ifndef N2V_BLK_MINB
#define N2V_BLK_MINB 5
#endif
// FULL: dim == 128 exactly (every lane holds 4 floats of every row, no masking)
template <bool ATOMIC, bool FULL>
__global__ void __launch_bounds__(SGNS_BLOCK, N2V_BLK_MINB)
sgns_group_kernel(TrainGroupsArgs a)
{
    constexpr int FN = BLK_FN;
    __shared__ float s_exp[EXP_TABLE_SIZE];
    __shared__ float4 s_orig[SGNS_BLOCK / 32][FN + 1][32];     // carried rows as first read
    for (int i = threadIdx.x; i < EXP_TABLE_SIZE; i += blockDim.x) s_exp[i] = exp_table_entry(i);
    __syncthreads();

    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = a.p.grid_warps;
    if (warp >= n_warps) return;
    const int32_t dim = FULL ? 128 : a.p.dim;
    const bool on = FULL || (lane * 4 < dim);
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const int32_t G = a.neg_group;
    const RowsFlat rows{a.syn0_part, a.syn1neg_part, dim};
    const int64_t first = a.dev_first ? *a.dev_first : a.first;
    int64_t end64 = a.dev_end ? *a.dev_end : a.end;
    if (end64 > a.capacity) end64 = a.capacity;
    if (first >= end64) return;
    // offsets below are relative to the stream's first word and 32-bit (the host checks the stream length)
    const uint32_t *const words = a.words + first;
    const uint32_t end = (uint32_t)(end64 - first);
    // this warp's contiguous range of the stream: the groups whose header lies in [lo, hi)
    const uint32_t per_warp = (uint32_t)(((uint64_t)end + (uint64_t)n_warps - 1) / (uint64_t)n_warps);
    const uint64_t lo64 = (uint64_t)warp * per_warp;
    if (lo64 >= end) return;
    const uint32_t lo = (uint32_t)lo64;
    const uint32_t hi = (end - lo > per_warp) ? lo + per_warp : end;
    uint32_t p = lo;
    for (;;) {                                                 // first header at or after lo
        const uint32_t wv = (p + lane < end) ? __ldcs(words + p + lane) : 0u;
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, (wv & GROUP_FLAG) != 0u);
        if (m) { p += __ffs(m) - 1; break; }
        if (hi - p <= 32u) return;
        p += 32;
    }
    if (p >= hi) return;

    uint32_t pairs = 0, carried = 0;                           // per warp and launch: far below 2^32
    auto sigmoid_g = [&](float f, float label, float alpha) -> float {
        return (label - s_exp[(int)((f + (float)MAX_EXP) * (float)(EXP_TABLE_SIZE / MAX_EXP / 2))]) * alpha;
    };
    auto load_tile = [&](uint32_t q) -> uint32_t {            // header + first 24 contexts of the group at q
        return (q < hi && q + lane < end) ? __ldcs(words + q + lane) : 0u;
    };
    auto poskey_of = [&](uint32_t h2) -> uint32_t { return G > 1 ? (h2 & 0xFFFFu) / (uint32_t)G : (h2 & 0xFFFFu); };

    uint32_t tile = load_tile(p);
    uint32_t pn = p + GROUP_HDR + (__shfl_sync(0xFFFFFFFFu, tile, 2) >> 16);
    uint32_t tile_n = load_tile(pn);
    // state of the carried negative set
    uint32_t set_s = 0xFFFFFFFFu, set_key = 0xFFFFFFFFu;
    bool set_live = false, set_dup = false;
    int32_t tg[FN];
    float4 out[FN + 1];
#pragma unroll
    for (int d = 0; d <= FN; ++d) out[d] = zero4;
#pragma unroll
    for (int d = 0; d < FN; ++d) tg[d] = 0;
    uint32_t alpha_s = 0xFFFFFFFFu; float alpha = 0.f;
    uint32_t set_hot = 0u;                                     // negatives of the live set that are not carried (hot rows)
    auto flush_set = [&]() {
        if (set_live && !set_dup) {
#pragma unroll
            for (int d = 1; d <= FN; ++d) {
                if ((set_hot >> d) & 1u) continue;
                const float4 og = s_orig[wib][d][lane];
                add_row<ATOMIC>(rows.r1(tg[d - 1]), lane,
                                make_float4(out[d].x - og.x, out[d].y - og.y, out[d].z - og.z, out[d].w - og.w), out[d], on);
            }
        }
        set_live = false;
    };

    while (p < hi) {
        const uint32_t h0 = __shfl_sync(0xFFFFFFFFu, tile, 0), h1 = __shfl_sync(0xFFFFFFFFu, tile, 1),
                       h2 = __shfl_sync(0xFFFFFFFFu, tile, 2);
        const int32_t centre = (int32_t)(h0 & ~GROUP_FLAG);
        const int32_t cnt = (int32_t)(h2 >> 16);
        const uint32_t key = poskey_of(h2);
        // the group after next: its tile is in flight during this group (its address needs tile_n's count)
        const uint32_t n1 = __shfl_sync(0xFFFFFFFFu, tile_n, 1), n2 = __shfl_sync(0xFFFFFFFFu, tile_n, 2);
        const bool have_next = pn < hi;
        const uint32_t pnn = pn + GROUP_HDR + (n2 >> 16);
        const uint32_t tile_nn = have_next ? load_tile(pnn) : 0u;
        if (h1 != alpha_s) { alpha = job_alpha(a.p, (int64_t)h1); alpha_s = h1; }
        auto ctx_at = [&](int32_t j) -> int32_t {
            return j < 32 - GROUP_HDR ? (int32_t)__shfl_sync(0xFFFFFFFFu, tile, GROUP_HDR + j)
                                      : (int32_t)__ldg(words + p + GROUP_HDR + j);
        };
        // ---- negative set of this group: the carried one is kept while the key is unchanged (G > 1) -- unless
        // this centre IS one of the carried rows: its row must then be read after the set's pending updates
        bool keep_set = set_live && h1 == set_s && key == set_key;
        bool collide = false;                      // (only ever true when G > 1 carries a set across centres)
        if (keep_set && !set_dup) {
#pragma unroll
            for (int d = 0; d < FN; ++d) collide |= (tg[d] == centre);
            keep_set = !collide;
        }
        if (!keep_set) {
            flush_set();
#pragma unroll
            for (int d = 0; d < FN; ++d) tg[d] = (int32_t)__shfl_sync(0xFFFFFFFFu, tile, 3 + d);
            set_dup = false;
#pragma unroll
            for (int d1 = 0; d1 < FN; ++d1)
#pragma unroll
                for (int d2 = d1 + 1; d2 < FN; ++d2) set_dup |= (tg[d1] == tg[d2]);
            set_hot = 0u;
            if (!set_dup) {
                // hot negatives (vocabulary row = local row * n_parts + part below hot_rows): reduced and re-read pair
                // by pair instead of carried (n2v_sgns.cu v3 says why)
#pragma unroll
                for (int d = 0; d < FN; ++d) if ((((int64_t)tg[d] << a.lg) | a.part) < a.p.hot_rows) set_hot |= 2u << d;
#pragma unroll
                for (int d = 0; d < FN; ++d) out[d + 1] = on ? ldcg4(rows.r1(tg[d]), lane) : zero4;
#pragma unroll
                for (int d = 1; d <= FN; ++d) s_orig[wib][d][lane] = out[d];
                carried += (uint32_t)FN;
            }
            set_live = true; set_s = h1; set_key = key;
        }
        const bool next_new = have_next && (n1 != set_s || poskey_of(n2) != set_key);
        if (!set_dup) {
            uint32_t skipmask = 0xC0u;                 // padding targets 6, 7
#pragma unroll
            for (int d = 0; d < FN; ++d) if (tg[d] == centre) skipmask |= 2u << d;     // skipped, not redrawn
            out[0] = on ? ldcg4(rows.r1(centre), lane) : zero4;
            s_orig[wib][0][lane] = out[0];
            ++carried;
            int32_t j = 0;
            int32_t ctx = ctx_at(0);
            float4 row1 = on ? ldcg4(rows.r0(ctx), lane) : zero4;
            const int32_t centre_n = (int32_t)(__shfl_sync(0xFFFFFFFFu, tile_n, 0) & ~GROUP_FLAG);
            if (have_next) {                           // next group's output rows: L2 warm-up
                if (on) prefetch_row_l2(rows.r1(centre_n), lane);
                if (next_new) {
#pragma unroll
                    for (int d = 0; d < FN; ++d) {
                        const int32_t tn = (int32_t)__shfl_sync(0xFFFFFFFFu, tile_n, 3 + d);
                        if (on) prefetch_row_l2(rows.r1(tn), lane);
                    }
                }
            }
            // two instantiations of the pair loop: sets with a hot row pay for the per-pair reductions / re-reads
            auto pair_loop = [&](auto hot_tag) {
            constexpr bool HOT = decltype(hot_tag)::value;
            while (j < cnt) {
                const int32_t jn = j + 1;
                // next input row always in flight (clamped past the group's end); it is stale only if it
                // is the very row this pair is about to update
                const int32_t ctx_n = jn < cnt ? ctx_at(jn) : ctx;
                const bool stale = ctx_n == ctx;
                const float4 row1n = on ? ldcg4(rows.r0(ctx_n), lane) : zero4;
                // 6 dot products by the transposing butterfly of the sentence-major kernel (n2v_sgns.cu)
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
                float gv = 0.0f;      // lane owns target lane >> 2 (0 = centre, 1..5 = negatives, 6,7 = padding)
                if (!((skipmask >> (lane >> 2)) & 1u) && fv > -(float)MAX_EXP && fv < (float)MAX_EXP)
                    gv = sigmoid_g(fv, lane < 4 ? 1.0f : 0.0f, alpha);
                float4 work = zero4;
#pragma unroll
                for (int d = 0; d <= FN; ++d) {       // g == 0: target skipped or |f| >= 6 (no-op)
                    const float gd = __shfl_sync(0xFFFFFFFFu, gv, d * 4);
                    axpy4(work, gd, out[d]);
                    axpy4(out[d], gd, row1);
                    if (HOT && d > 0 && ((set_hot >> d) & 1u) && !((skipmask >> d) & 1u)) {   // hot row: update now, re-read
                        float *const rp = rows.r1(tg[d - 1]);
                        add_row<ATOMIC>(rp, lane, make_float4(gd * row1.x, gd * row1.y, gd * row1.z, gd * row1.w), out[d], on);
                        if (ATOMIC) out[d] = on ? ldcg4(rp, lane) : zero4;
                    }
                }
                float4 upd1 = row1;
                upd1.x += work.x; upd1.y += work.y; upd1.z += work.z; upd1.w += work.w;
                add_row<ATOMIC>(rows.r0(ctx), lane, work, upd1, on);
                j = jn;
                row1 = row1n;
                if (stale && j < cnt) row1 = on ? ldcg4(rows.r0(ctx), lane) : zero4;   // re-read after the update
                ctx = ctx_n;
            }
            };
            if (set_hot) pair_loop(std::true_type{}); else pair_loop(std::false_type{});
            {                                          // the centre row: one reduction of what this group added
                const float4 og = s_orig[wib][0][lane];
                add_row<ATOMIC>(rows.r1(centre), lane,
                                make_float4(out[0].x - og.x, out[0].y - og.y, out[0].z - og.z, out[0].w - og.w), out[0], on);
            }
            if (G > 1) {                               // a carried negative that IS this centre is stale now: drop the set
                bool again = false;
#pragma unroll
                for (int d = 0; d < FN; ++d) again |= (tg[d] == centre);
                if (again) flush_set();
            }
        } else {
            // repeated row in the set: uncarried sequential form (every target re-read per pair), gensim's
            // semantics for a repeated draw -- the sentence-major kernel's fallback
            const bool act1[1] = {on};
            const int32_t t_lanes = lane == 0 ? tg[0] : lane == 1 ? tg[1] : lane == 2 ? tg[2] : lane == 3 ? tg[3] : tg[4];
            for (int32_t j = 0; j < cnt; ++j)
                train_pair<1, ATOMIC>(rows, dim, centre, ctx_at(j), t_lanes, FN, alpha, act1, s_exp, lane);
        }
        pairs += (uint32_t)cnt;
        p = pn; pn = pnn;
        tile = tile_n; tile_n = tile_nn;
    }
    flush_set();
    if (lane == 0 && a.pairs_out && pairs) {
        atomicAdd(a.pairs_out, (unsigned long long)pairs); atomicAdd(a.pairs_out + 1, (unsigned long long)carried);
    }
}

static inline size_t blk_align(size_t x, size_t al = 256) { return (x + al - 1) / al * al; }

static int log2_parts(int32_t n_parts)
{
    return n_parts == 1 ? 0 : n_parts == 2 ? 1 : n_parts == 4 ? 2 : n_parts == 8 ? 3 : -1;
}

// per-(stream, sentence) word counts are int32, their running sum is not (8 parts x 4 M walks: > 2^31 words)
struct WidenCount { __host__ __device__ int64_t operator()(int32_t c) const { return (int64_t)c; } };
using WideCounts = cub::TransformInputIterator<int64_t, WidenCount, const int32_t *>;

static size_t groups_scan_bytes(int64_t n)
{
    size_t b = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, b, WideCounts((const int32_t *)nullptr, WidenCount()), (int64_t *)nullptr, n);
    return blk_align(b);
}

static int groups_args(GroupsArgs &g, const int32_t *tokens, const int64_t *sent_off, int64_t n_sent, int32_t stride,
                       int64_t sent_id_base, const int32_t *vocab_of_id, const uint32_t *keep_thr,
                       const n2v_sgns_params_t *params, int32_t part, int32_t n_parts)
{
    N2V_REQUIRE(params, "params is NULL");
    N2V_REQUIRE(tokens && n_sent >= 0 && n_sent < 2147483647ll, "bad corpus");
    N2V_REQUIRE(sent_off || stride > 0, "sent_off is NULL and stride <= 0");
    N2V_REQUIRE(params->window >= 1 && params->window <= SGNS_MAX_WINDOW, "window out of range (1..96)");
    N2V_REQUIRE(params->max_sentence_len >= 1 && params->max_sentence_len <= 65535, "max_sentence_len out of range");
    const int lg = log2_parts(n_parts);
    N2V_REQUIRE(lg >= 0 && part >= 0 && part < n_parts, "n_parts must be 1, 2, 4 or 8 and 0 <= part < n_parts");
    memset(&g, 0, sizeof(g));
    g.a.tokens = tokens; g.a.sent_off = sent_off; g.a.n_sent = n_sent; g.a.stride = stride;
    g.a.sent_id_base = sent_id_base; g.a.vocab_of_id = vocab_of_id; g.a.keep_thr = keep_thr; g.a.p = *params;
    g.part = part; g.lg = lg; g.n_parts = n_parts;
    return N2V_OK;
}

}  // namespace n2v

using namespace n2v;

extern "C" size_t n2v_sgns_groups_workspace_bytes(int64_t n_sent, int32_t n_parts)
{
    const int64_t n = (int64_t)n_parts * (n_sent > 0 ? n_sent : 0) + 1;
    return blk_align(sizeof(int32_t) * (size_t)n) + groups_scan_bytes(n);
}

extern "C" int n2v_sgns_groups_count(const int32_t *tokens, const int64_t *sent_off, int64_t n_sent, int32_t stride,
                                     int64_t sent_id_base, const int32_t *vocab_of_id, const uint32_t *keep_thr,
                                     const n2v_sgns_params_t *params, int32_t part, int32_t n_parts,
                                     int64_t *offsets, void *workspace, size_t workspace_bytes, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    GroupsArgs g;
    int rc = groups_args(g, tokens, sent_off, n_sent, stride, sent_id_base, vocab_of_id, keep_thr, params, part, n_parts);
    if (rc != N2V_OK) return rc;
    N2V_REQUIRE(offsets && workspace, "offsets / workspace is NULL");
    const int64_t n = (int64_t)n_parts * n_sent + 1;
    N2V_REQUIRE(workspace_bytes >= n2v_sgns_groups_workspace_bytes(n_sent, n_parts), "groups workspace too small");
    g.counts = (int32_t *)workspace;
    void *tmp = (char *)workspace + blk_align(sizeof(int32_t) * (size_t)n);
    size_t tmp_bytes = groups_scan_bytes(n);
    N2V_CHECK_CUDA(cudaMemsetAsync(g.counts + (n - 1), 0, sizeof(int32_t), stream));
    if (n_sent > 0) {
        int sms = sm_count();
        if (sms <= 0) { set_error("no CUDA device"); return N2V_ECUDA; }
        int64_t blocks = (n_sent + SGNS_BLOCK / 32 - 1) / (SGNS_BLOCK / 32);
        if (blocks > (int64_t)sms * 16) blocks = (int64_t)sms * 16;
        sgns_groups_kernel<false><<<(unsigned)blocks, SGNS_BLOCK, 0, stream>>>(g);
        N2V_LAUNCH_CHECK();
    }
    N2V_CHECK_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, WideCounts((const int32_t *)g.counts, WidenCount()), offsets, n,
                                                 stream));
    return N2V_OK;
}

extern "C" int n2v_sgns_groups_fill(const int32_t *tokens, const int64_t *sent_off, int64_t n_sent, int32_t stride,
                                    int64_t sent_id_base, const int32_t *vocab_of_id, const uint32_t *keep_thr,
                                    const n2v_sgns_params_t *params, int32_t part, int32_t n_parts,
                                    const uint32_t *cum_table, const int32_t *bucket_lo, int32_t neg_group,
                                    const int64_t *offsets, uint32_t *words, int64_t capacity_words,
                                    unsigned long long *overflow, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    GroupsArgs g;
    int rc = groups_args(g, tokens, sent_off, n_sent, stride, sent_id_base, vocab_of_id, keep_thr, params, part, n_parts);
    if (rc != N2V_OK) return rc;
    N2V_REQUIRE(offsets && words && overflow && capacity_words >= 0, "offsets / words / overflow is NULL");
    N2V_REQUIRE(cum_table && bucket_lo, "cum_table / bucket_lo is NULL");
    N2V_REQUIRE(neg_group >= 1 && neg_group <= 256, "neg_group must be in [1, 256]");
    N2V_REQUIRE(params->V >= n_parts && params->bucket_bits >= 1 && params->bucket_bits <= 24, "bad vocabulary parameters");
    if (n_sent == 0) return N2V_OK;
    g.offsets = offsets; g.words = words; g.capacity = capacity_words; g.overflow = overflow;
    g.cum_table = cum_table; g.bucket_lo = bucket_lo; g.neg_group = neg_group;
    int sms = sm_count();
    if (sms <= 0) { set_error("no CUDA device"); return N2V_ECUDA; }
    int64_t blocks = (n_sent + SGNS_BLOCK / 32 - 1) / (SGNS_BLOCK / 32);
    if (blocks > (int64_t)sms * 16) blocks = (int64_t)sms * 16;
    sgns_groups_kernel<true><<<(unsigned)blocks, SGNS_BLOCK, 0, stream>>>(g);
    N2V_LAUNCH_CHECK();
    return N2V_OK;
}

extern "C" int n2v_sgns_train_groups(const uint32_t *words, int64_t first_word, int64_t n_words,
                                     const int64_t *dev_first, const int64_t *dev_end, int64_t capacity_words,
                                     const n2v_sgns_params_t *params, int32_t neg_group, float *syn0_part,
                                     float *syn1neg_part, int32_t part, int32_t n_parts,
                                     unsigned long long *pairs_out, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    N2V_REQUIRE(params, "params is NULL");
    N2V_REQUIRE(first_word >= 0 && n_words >= 0 && capacity_words >= 0, "negative size");
    N2V_REQUIRE(n_words < 4294967296ll - 64, "a stream holds at most 2^32 words per launch");
    N2V_REQUIRE((dev_first == nullptr) == (dev_end == nullptr), "dev_first and dev_end go together");
    if (!dev_first && n_words == 0) return N2V_OK;
    N2V_REQUIRE(words && syn0_part && syn1neg_part, "NULL buffer");
    const int lg = log2_parts(n_parts);
    N2V_REQUIRE(lg >= 0 && part >= 0 && part < n_parts, "n_parts must be 1, 2, 4 or 8 and 0 <= part < n_parts");
    const n2v_sgns_params_t &p = *params;
    N2V_REQUIRE(p.V >= n_parts, "fewer vocabulary rows than parts");
    N2V_REQUIRE(p.dim >= 4 && p.dim <= 128 && p.dim % 4 == 0, "block kernel: dim must be a multiple of 4, <= 128");
    N2V_REQUIRE(p.negative == BLK_FN, "block kernel: negative must be 5");
    N2V_REQUIRE(neg_group >= 1 && neg_group <= 256, "neg_group must be in [1, 256]");
    N2V_REQUIRE(p.grid_warps >= 1 && p.total_examples >= 1 && p.sent_per_job >= 1, "bad schedule");
    if (sm_count() <= 0) { set_error("no CUDA device"); return N2V_ECUDA; }
    TrainGroupsArgs a;
    memset(&a, 0, sizeof(a));
    a.words = words; a.first = first_word; a.end = first_word + n_words; a.capacity = capacity_words;
    a.dev_first = dev_first; a.dev_end = dev_end;
    a.syn0_part = syn0_part; a.syn1neg_part = syn1neg_part;
    a.p = p; a.part = part; a.lg = lg; a.neg_group = neg_group; a.pairs_out = pairs_out;
    if (!dev_first) {                      // no more warps than groups could exist (a group is >= 9 words)
        const int64_t most = (n_words + GROUP_HDR) / (GROUP_HDR + 1);
        if (most < a.p.grid_warps) a.p.grid_warps = (int32_t)(most > 0 ? most : 1);
    }
    const int wpb = SGNS_BLOCK / 32;
    const int blocks = (a.p.grid_warps + wpb - 1) / wpb;
    const bool full = p.dim == 128;
    if (p.atomic_updates) {
        if (full) sgns_group_kernel<true, true><<<blocks, SGNS_BLOCK, 0, stream>>>(a);
        else sgns_group_kernel<true, false><<<blocks, SGNS_BLOCK, 0, stream>>>(a);
    } else {
        if (full) sgns_group_kernel<false, true><<<blocks, SGNS_BLOCK, 0, stream>>>(a);
        else sgns_group_kernel<false, false><<<blocks, SGNS_BLOCK, 0, stream>>>(a);
    }
    N2V_LAUNCH_CHECK();
    return N2V_OK;
}
