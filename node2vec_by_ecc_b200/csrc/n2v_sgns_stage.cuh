// Device helpers shared by the SGNS kernels (n2v_sgns.cu: sentence-major kernels; n2v_sgns_block.cu:
// the block-partitioned pair kernels): gensim constants, the sigmoid table entry, the count^0.75 draw,
// the per-warp sentence staging (sub-sample + window shrink, train_batch_sg prologue) and row helpers.
#pragma once
#include "n2v_common.cuh"

namespace n2v {

constexpr int EXP_TABLE_SIZE = 1000;
constexpr int MAX_EXP = 6;
constexpr int SGNS_BLOCK = 128;          // 4 warps
constexpr int SGNS_MAX_NEG = 16;
constexpr int SGNS_SMEM_TOKENS = 256;    // per-warp staging of the kept tokens of a sentence chunk
// next_chunk carries 2 * window kept tokens between chunks and load_chunk only appends while at most
// SGNS_SMEM_TOKENS - 32 are staged: a larger window would never make progress
constexpr int SGNS_MAX_WINDOW = (SGNS_SMEM_TOKENS - 32 - 32) / 2;   // 96

// word2vec_inner.pyx init(): EXP_TABLE[i] = exp((i / 1000 * 2 - 1) * 6); then x / (x + 1), with the
// Cython code's float32 casts (float argument, C double exp, float result, float division).
// Every block rebuilds its shared-memory copy: 8 exps per thread, no global state, no host sync.
__device__ __forceinline__ float exp_table_entry(int i)
{
    const float x = __fmul_rn(__fadd_rn(__fmul_rn(__fdiv_rn((float)i, (float)EXP_TABLE_SIZE), 2.0f), -1.0f), (float)MAX_EXP);
    const float e = (float)exp((double)x);
    return __fdiv_rn(e, __fadd_rn(e, 1.0f));
}

__device__ __forceinline__ int32_t draw_negative(uint32_t r32, const uint32_t *__restrict__ cum_table,
                                                 const int32_t *__restrict__ bucket_lo, int32_t V,
                                                 int32_t bucket_bits)
{
    // bisect_left(cum_table, (next_random >> 16) % cum_table[-1])
    const uint32_t cum_last = __ldg(cum_table + V - 1);
    const uint32_t r = r32 % cum_last;
    const uint32_t b = r >> (31 - bucket_bits);
    int32_t lo = __ldg(bucket_lo + b), hi = __ldg(bucket_lo + b + 1);
    while (lo < hi) {
        int32_t mid = (lo + hi) >> 1;
        if (__ldg(cum_table + mid) < r) lo = mid + 1; else hi = mid;
    }
    return lo;
}

struct SgnsArgs {
    const int32_t *tokens; const int64_t *sent_off; int64_t n_sent; int32_t stride;
    int64_t sent_id_base; const int32_t *vocab_of_id;
    const uint32_t *keep_thr; const uint32_t *cum_table; const int32_t *bucket_lo;
    n2v_sgns_params_t p;
    float *syn0, *syn1neg; unsigned long long *pairs_out;
    float *parts0[8], *parts1[8]; int32_t parts_log2;     // sharded tables (v3 only): n_parts = 1 << parts_log2
};

struct WarpSentence {       // per-warp staging of the kept tokens of one sentence chunk
    int32_t *idx; uint16_t *pos; uint8_t *rw;
};

// job_producer: alpha is fixed per job of whole sentences (word2vec.py train())
__device__ __forceinline__ float job_alpha(const n2v_sgns_params_t &p, int64_t s)
{
    const int64_t ex = p.example_base + s;
    const int64_t job_first = ex - ex % p.sent_per_job;
    double prog = (double)job_first / (double)p.total_examples;
    double al = (double)p.alpha0 - ((double)p.alpha0 - (double)p.min_alpha) * prog;
    return (float)(al > (double)p.min_alpha ? al : (double)p.min_alpha);
}

// train_batch_sg prologue: sub-sample + per-position window shrink, compacted in sentence order.
// Appends to the n_kept tokens already staged from tokens [t_next, tl) until the buffer is full;
// returns the new n_kept.
__device__ __forceinline__ int32_t load_chunk(const SgnsArgs &a, const WarpSentence &ws, int64_t tb, int64_t tl,
                                              int64_t &t_next, uint64_t gs, uint32_t ep8, uint32_t k0,
                                              uint32_t k1, int lane, int32_t n_kept)
{
    while (t_next < tl && n_kept <= SGNS_SMEM_TOKENS - 32) {
        const int64_t t = t_next + lane;
        int32_t wv = -1; uint32_t red = 0;
        if (t < tl) {
            int32_t id = a.tokens[tb + t];
            if (id >= 0) wv = a.vocab_of_id ? __ldg(a.vocab_of_id + id) : id;
            if (wv >= 0) {
                const Philox4 r = philox4x32_10((uint32_t)gs, (uint32_t)(gs >> 32), (uint32_t)t, ep8, k0, k1);
                if (a.keep_thr && __ldg(a.keep_thr + wv) < r.x) wv = -1;   // sample_int < random_int32
                red = r.y % (uint32_t)a.p.window;                          // reduced_windows[i]
            }
        }
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, wv >= 0);
        if (wv >= 0) {
            const int o = n_kept + __popc(m & ((1u << lane) - 1u));
            ws.idx[o] = wv; ws.pos[o] = (uint16_t)t; ws.rw[o] = (uint8_t)red;
        }
        n_kept += __popc(m);
        t_next += 32;
    }
    __syncwarp();
    return n_kept;
}

// Streams a sentence of any length (<= max_sentence_len) through the per-warp staging buffer with
// exact windows: centres [c_lo, c_hi) of the buffer are the ones whose full window is present;
// between chunks the last 2*window kept tokens are carried over (window of left context + the
// window of centres that still lacked their right context). Returns false when the sentence is done.
__device__ __forceinline__ bool next_chunk(const SgnsArgs &a, const WarpSentence &ws, int64_t tb, int64_t tl,
                                           int64_t &t_next, uint64_t gs, uint32_t ep8, uint32_t k0, uint32_t k1,
                                           int lane, int32_t &n_kept, int32_t &c_lo, int32_t &c_hi, bool &first)
{
    const int32_t window = a.p.window;
    if (first) { first = false; n_kept = 0; c_lo = 0; }
    else {
        if (t_next >= tl) return false;                     // the previous chunk was the last one
        const int32_t src = c_hi - window, keep = n_kept - src;   // = 2 * window
        for (int32_t c = 0; c < keep; c += 32) {
            const int32_t k = c + lane;
            int32_t vi = 0; uint16_t vp = 0; uint8_t vr = 0;
            if (k < keep) { vi = ws.idx[src + k]; vp = ws.pos[src + k]; vr = ws.rw[src + k]; }
            __syncwarp();
            if (k < keep) { ws.idx[k] = vi; ws.pos[k] = vp; ws.rw[k] = vr; }
            __syncwarp();
        }
        n_kept = keep; c_lo = window;
    }
    n_kept = load_chunk(a, ws, tb, tl, t_next, gs, ep8, k0, k1, lane, n_kept);
    c_hi = (t_next >= tl) ? n_kept : n_kept - window;
    return true;
}

__device__ __forceinline__ float4 ldcg4(const float *row, int lane)
{
    return __ldcg(reinterpret_cast<const float4 *>(row) + lane);
}
// pull a row's line(s) towards L2 without occupying registers or ordering against later accesses
__device__ __forceinline__ void prefetch_row_l2(const float *row, int lane)
{
    asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const float4 *>(row) + lane));
}
__device__ __forceinline__ float dot4(const float4 &x, const float4 &y)
{
    return x.x * y.x + x.y * y.y + x.z * y.z + x.w * y.w;
}
__device__ __forceinline__ void axpy4(float4 &acc, float g, const float4 &x)
{
    acc.x += g * x.x; acc.y += g * x.y; acc.z += g * x.z; acc.w += g * x.w;
}
template <bool ATOMIC>
__device__ __forceinline__ void add_row(float *row, int lane, const float4 &delta, const float4 &updated, bool on)
{
    if (!on) return;
    float4 *p = reinterpret_cast<float4 *>(row) + lane;
    if (ATOMIC) atomicAdd(p, delta); else *p = updated;
}

}  // namespace n2v
