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
// Packed fp32 pairs (sm_100a FFMA2 / FMUL2: one issue slot for two fused multiply-adds). A float4 from a
// 128-bit load sits in an aligned register quad, so (x, y) and (z, w) are register pairs and the
// mov.b64 packs below cost nothing. Each half is an ordinary fma.rn, so the per-element results are those
// of scalar code; only the dot product's summation order differs: (x.x*y.x + x.z*y.z) + (x.y*y.y + x.w*y.w).
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t pack2(float a, float b)
{
    f32x2_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r;
}
__device__ __forceinline__ void unpack2(f32x2_t v, float &a, float &b)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ f32x2_t fma2(f32x2_t a, f32x2_t b, f32x2_t c)
{
    f32x2_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;
}
__device__ __forceinline__ f32x2_t mul2(f32x2_t a, f32x2_t b)
{
    f32x2_t d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}
__device__ __forceinline__ float dot4(const float4 &x, const float4 &y)
{
    f32x2_t t = mul2(pack2(x.x, x.y), pack2(y.x, y.y));
    t = fma2(pack2(x.z, x.w), pack2(y.z, y.w), t);
    float a, b;
    unpack2(t, a, b);
    return a + b;
}
__device__ __forceinline__ void axpy4(float4 &acc, float g, const float4 &x)
{
    const f32x2_t gg = pack2(g, g);
    const f32x2_t lo = fma2(gg, pack2(x.x, x.y), pack2(acc.x, acc.y));
    const f32x2_t hi = fma2(gg, pack2(x.z, x.w), pack2(acc.z, acc.w));
    unpack2(lo, acc.x, acc.y);
    unpack2(hi, acc.z, acc.w);
}
template <bool ATOMIC>
__device__ __forceinline__ void add_row(float *row, int lane, const float4 &delta, const float4 &updated, bool on)
{
    if (!on) return;
    float4 *p = reinterpret_cast<float4 *>(row) + lane;
    if (ATOMIC) atomicAdd(p, delta); else *p = updated;
}

// ---- training -------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, o);
    return v;
}

// One (centre, context) pair == one fast_sentence_sg_neg call. my_t = this lane's negative draw
// (lane n holds negative n). Rows: row1 = syn0[context] (input), row2 = syn1neg[target].
template <int NV, bool ATOMIC>
__device__ __forceinline__ void apply_target(float4 *row2p, float4 (&row2)[NV], const float4 (&row1)[NV],
                                             float4 (&work)[NV], float g, const bool (&act)[NV], int lane)
{
#pragma unroll
    for (int c = 0; c < NV; ++c) {
        work[c].x += g * row2[c].x; work[c].y += g * row2[c].y;       // work += g * syn1neg[t]
        work[c].z += g * row2[c].z; work[c].w += g * row2[c].w;
        if (act[c]) {                                                 // syn1neg[t] += g * row1
            if (ATOMIC) {
                atomicAdd(row2p + c * 32 + lane,
                          make_float4(g * row1[c].x, g * row1[c].y, g * row1[c].z, g * row1[c].w));
            } else {
                row2[c].x += g * row1[c].x; row2[c].y += g * row1[c].y;
                row2[c].z += g * row1[c].z; row2[c].w += g * row1[c].w;
                row2p[c * 32 + lane] = row2[c];
            }
        }
    }
}

// Where the rows live. Flat: one [V, dim] table per side on this device. Sharded: row i of a table
// lives in part i % n_parts (n_parts a power of two; parts may be peer-GPU memory mapped over
// NVLink) at local row i / n_parts -- vocabulary order is count-descending, so the parts carry
// equal shares of the traffic.
struct RowsFlat {
    float *s0, *s1; int32_t dim;
    __device__ __forceinline__ float *r0(int32_t i) const { return s0 + (int64_t)i * dim; }
    __device__ __forceinline__ float *r1(int32_t i) const { return s1 + (int64_t)i * dim; }
};
struct RowsSharded {
    float *const *p0; float *const *p1; int32_t dim, lg, mask;      // p0/p1: 2 x n_parts pointers in shared memory
    __device__ __forceinline__ float *r0(int32_t i) const { return p0[i & mask] + (int64_t)(i >> lg) * dim; }
    __device__ __forceinline__ float *r1(int32_t i) const { return p1[i & mask] + (int64_t)(i >> lg) * dim; }
};

template <int NV, bool ATOMIC, class Rows>
__device__ __forceinline__ void train_pair(const Rows rows, int32_t dim,
                                           int32_t centre, int32_t ctx, int32_t my_t, int32_t negative,
                                           float alpha, const bool (&act)[NV], const float *s_exp, int lane)
{
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 *row1p = reinterpret_cast<float4 *>(rows.r0(ctx));
    float4 row1[NV], work[NV];
#pragma unroll
    for (int c = 0; c < NV; ++c) { row1[c] = act[c] ? row1p[c * 32 + lane] : zero4; work[c] = zero4; }

    constexpr int FN = 5;   // fast path: gensim's default negative=5, dim <= 128
    bool fast = (NV == 1) && (negative == FN);
    int32_t tg[FN + 1];
    if (fast) {
        tg[0] = centre;
#pragma unroll
        for (int d = 1; d <= FN; ++d) tg[d] = __shfl_sync(0xFFFFFFFFu, my_t, d - 1);
        // a repeated negative must see the row as updated by its first occurrence: slow path
#pragma unroll
        for (int d1 = 1; d1 <= FN; ++d1)
#pragma unroll
            for (int d2 = d1 + 1; d2 <= FN; ++d2) if (tg[d1] == tg[d2]) fast = false;
    }
    if (fast) {
        // all target rows in flight at once: one HBM latency per pair instead of six
        float4 r2[FN + 1][1];
        float f[FN + 1];
#pragma unroll
        for (int d = 0; d <= FN; ++d) {
            const float4 *rp = reinterpret_cast<const float4 *>(rows.r1(tg[d]));
            r2[d][0] = (act[0] && !(d > 0 && tg[d] == centre)) ? rp[lane] : zero4;
        }
#pragma unroll
        for (int d = 0; d <= FN; ++d)
            f[d] = row1[0].x * r2[d][0].x + row1[0].y * r2[d][0].y + row1[0].z * r2[d][0].z + row1[0].w * r2[d][0].w;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int d = 0; d <= FN; ++d) f[d] += __shfl_xor_sync(0xFFFFFFFFu, f[d], o);
#pragma unroll
        for (int d = 0; d <= FN; ++d) {
            if (d > 0 && tg[d] == centre) continue;                               // skipped, not redrawn
            if (f[d] <= -(float)MAX_EXP || f[d] >= (float)MAX_EXP) continue;
            const float sg = s_exp[(int)((f[d] + (float)MAX_EXP) * (float)(EXP_TABLE_SIZE / MAX_EXP / 2))];
            const float g = ((d == 0 ? 1.0f : 0.0f) - sg) * alpha;
            float4 *row2p = reinterpret_cast<float4 *>(rows.r1(tg[d]));
            apply_target<1, ATOMIC>(row2p, r2[d], reinterpret_cast<const float4 (&)[1]>(row1),
                                    reinterpret_cast<float4 (&)[1]>(work), g,
                                    reinterpret_cast<const bool (&)[1]>(act), lane);
        }
    } else {
        for (int32_t d = 0; d <= negative; ++d) {
            int32_t target; float label;
            if (d == 0) { target = centre; label = 1.0f; }
            else {
                target = __shfl_sync(0xFFFFFFFFu, my_t, d - 1);
                if (target == centre) continue;
                label = 0.0f;
            }
            float4 *row2p = reinterpret_cast<float4 *>(rows.r1(target));
            float4 row2[NV];
            float f = 0.0f;
#pragma unroll
            for (int c = 0; c < NV; ++c) {
                row2[c] = act[c] ? row2p[c * 32 + lane] : zero4;
                f += row1[c].x * row2[c].x + row1[c].y * row2[c].y + row1[c].z * row2[c].z + row1[c].w * row2[c].w;
            }
            f = warp_sum(f);
            if (f <= -(float)MAX_EXP || f >= (float)MAX_EXP) continue;
            const float sg = s_exp[(int)((f + (float)MAX_EXP) * (float)(EXP_TABLE_SIZE / MAX_EXP / 2))];
            const float g = (label - sg) * alpha;
            apply_target<NV, ATOMIC>(row2p, row2, row1, work, g, act, lane);
        }
    }
#pragma unroll
    for (int c = 0; c < NV; ++c) {
        if (act[c]) {                                                             // syn0[ctx] += work
            if (ATOMIC) atomicAdd(row1p + c * 32 + lane, work[c]);
            else {
                row1[c].x += work[c].x; row1[c].y += work[c].y;
                row1[c].z += work[c].z; row1[c].w += work[c].w;
                row1p[c * 32 + lane] = row1[c];
            }
        }
    }
}


}  // namespace n2v
