// Rejection-sampling walker, second form: hashed distance-1 test + per-lane state machine.
//
// What the first form (n2v_walk.cu: walk_reject_kernel) pays for, measured with ncu on C4
// (profiles/r01_a_walk_reject_ncu_full.json): 31 DRAM sectors per step for 127 algorithmic bytes,
// because every probe of the binary search in adj(prev) pulls a 32-byte sector for 4 bytes and
// prev is usually a hub (log2(deg) ~ 11-17 probes); and a warp that advances step by step in
// lock-step waits for its slowest lane at every step (trials are geometric: the max over 32 lanes
// is ~4x the mean).
//
// Here (a) the distance-1 test G.has_edge(x, prev) (node2vec.py:144) is ONE lookup in an
// open-addressing hash set of all arcs (key = x<<32 | prev, load factor <= 0.5, linear probing:
// ~1.3 slots, consecutive slots share a sector), and (b) every lane runs its own walker as a state
// machine -- ROW (fetch packed row start/degree of cur), [ALIAS (weighted: node-table slot)],
// COL (fetch the candidate), PROBE (hash slots) -- and each loop iteration issues exactly one
// 8-byte load per lane whatever its state, so lanes never wait for each other's trials.
// Same transition law, same Philox addressing (walk_id, step, trial) as the first form.
#include "n2v_common.cuh"
#include "n2v_reject.cuh"

namespace n2v {

constexpr int W2_BLOCK = 256;
constexpr uint64_t HASH_EMPTY = 0xFFFFFFFFFFFFFFFFull;
constexpr int PACK_DEG_BITS = 24;

__global__ void pack_rows_kernel(const int64_t *__restrict__ row_ptr, int32_t n, uint64_t *__restrict__ packed,
                                 int *__restrict__ overflow)
{
    int32_t v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    const int64_t b = row_ptr[v], d = row_ptr[v + 1] - b;
    if (d >= (1ll << PACK_DEG_BITS) || b >= (1ll << (64 - PACK_DEG_BITS))) { *overflow = 1; }
    packed[v] = ((uint64_t)b << PACK_DEG_BITS) | (uint64_t)(d & ((1ll << PACK_DEG_BITS) - 1));
}

__global__ void edge_hash_build_kernel(const int64_t *__restrict__ row_ptr, const int32_t *__restrict__ col,
                                       int32_t n, int64_t nnz, unsigned long long *__restrict__ table, uint64_t mask)
{
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= nnz) return;
    int32_t lo = 0, hi = n;                       // row of arc e
    while (hi - lo > 1) {
        int32_t mid = lo + ((hi - lo) >> 1);
        if (__ldg(row_ptr + mid) <= e) lo = mid; else hi = mid;
    }
    const uint64_t key = ((uint64_t)(uint32_t)lo << 32) | (uint32_t)col[e];
    uint64_t slot = mix64(key) & mask;
    for (;;) {
        unsigned long long old = atomicCAS(table + slot, (unsigned long long)HASH_EMPTY, (unsigned long long)key);
        if (old == HASH_EMPTY || old == key) return;
        slot = (slot + 1) & mask;
    }
}

enum : int { ST_ROW = 0, ST_ALIAS = 1, ST_COL = 2, ST_PROBE = 3, ST_DONE = 4, ST_STR = 5, ST_WRET = 6 };

#ifndef N2V_W2_MINB
#define N2V_W2_MINB 5
#endif
template <bool WEIGHTED>
__global__ void __launch_bounds__(W2_BLOCK, N2V_W2_MINB)
walk_reject_indexed_kernel(const uint64_t *__restrict__ packed_rows, const int32_t *__restrict__ col,
                           const n2v_slot_t *__restrict__ node_slots, const n2v_slot_t *__restrict__ first_slots,
                           const double *__restrict__ w, const double *__restrict__ strength,
                           const unsigned long long *__restrict__ edge_hash, uint64_t hash_mask, RejectParams rp,
                           int64_t nnz, const int32_t *__restrict__ starts, int64_t n_walks, int32_t L, uint32_t k0,
                           uint32_t k1, uint64_t walk_id_base, int32_t *__restrict__ walks,
                           int32_t *__restrict__ lens, unsigned long long *__restrict__ counters)
{
    __shared__ int32_t stage_all[W2_BLOCK][8 + 1];
    int32_t *stage = stage_all[threadIdx.x];
    const int lane = threadIdx.x & 31;
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const bool live = i < n_walks && L > 0;
    const uint64_t wid = walk_id_base + (uint64_t)i;
    int32_t *const out = walks + i * (int64_t)L;

    int32_t cur = live ? starts[i] : 0, prev = -1;
    int32_t len = live ? 1 : 0;                   // tokens produced so far == index of the next token
    int64_t b = 0; int32_t K = 0;                 // row of cur
    int32_t pdeg = 0;                             // degree of prev (charged probes)
    uint32_t trial = 0, y_acc = 0, y_al = 0; int32_t x = -1; int64_t kk = 0; uint64_t slot = 0, key = 0;
    double w_ret = 1.0, w_row = 0.0;              // weighted fold: weight of (prev,cur), total weight of cur's row
    const bool wfold = WEIGHTED && rp.fold;
    int state = (live && L > 1) ? ST_ROW : ST_DONE;
    uint32_t n_trials = 0, n_tests = 0, n_probes = 0;   // per walker: < 2^32
    if (live) stage[0] = cur;

    auto emit = [&](int32_t tok) {                // append a token; one full 32-byte sector per 8 tokens
        stage[len & 7] = tok;
        ++len;
        if ((len & 7) == 0) {
#pragma unroll
            for (int c = 0; c < 8; ++c) out[len - 8 + c] = stage[c];
        }
    };
    auto finish = [&]() {                         // write the tail, pad with -1 (dead end / end of walk)
        for (int32_t t = len & ~7; t < L; ++t) out[t] = t < len ? stage[t & 7] : -1;
        state = ST_DONE;
    };
    if (live && L == 1) out[0] = cur;

    while (__any_sync(0xFFFFFFFFu, state != ST_DONE)) {
        // ---- exactly one 8-byte load per lane, address by state
        unsigned long long v = 0;
        if (state == ST_ROW) v = __ldg(reinterpret_cast<const unsigned long long *>(packed_rows + cur));
        else if (state == ST_ALIAS)       // the first step may have its own table (popularity node tables)
            v = __ldg(reinterpret_cast<const unsigned long long *>(((prev < 0 && first_slots) ? first_slots : node_slots) + b + kk));
        else if (state == ST_STR) v = __ldg(reinterpret_cast<const unsigned long long *>(strength + cur));
        else if (state == ST_WRET) v = __ldg(reinterpret_cast<const unsigned long long *>(w + slot));
        else if (state == ST_COL) {
            const int64_t e = b + kk, e2 = e & ~1ll;
            if (e2 + 1 < nnz) v = __ldg(reinterpret_cast<const unsigned long long *>(col + e2));
            else v = (unsigned long long)(uint32_t)__ldg(col + e) << ((e & 1) ? 32 : 0);
        } else if (state == ST_PROBE) v = __ldg(edge_hash + slot);

        // ---- consume
        bool accept = false, draw = false, outlier = false;
        if (state == ST_ROW) {
            b = (int64_t)(v >> PACK_DEG_BITS); K = (int32_t)(v & ((1ull << PACK_DEG_BITS) - 1));
            if (K <= 0) finish();                 // dead end (node2vec.py:76-77)
            else if (wfold && prev >= 0) state = ST_STR;
            else { trial = 0; draw = true; }
        } else if (state == ST_STR) {
            w_row = __longlong_as_double((long long)v); trial = 0; draw = true;
        } else if (state == ST_WRET) {
            w_ret = __longlong_as_double((long long)v); state = ST_ROW;
        } else if (state == ST_ALIAS) {           // static law ~ w(cur, .): the node alias table
            if (!(y_al < (uint32_t)(v >> 32))) kk = (int64_t)(int32_t)(uint32_t)v;
            state = ST_COL;
        } else if (state == ST_COL) {
            x = (int32_t)(((b + kk) & 1) ? (uint32_t)(v >> 32) : (uint32_t)v);
            if (prev < 0) accept = true;                       // first step: one static draw (:69-70)
            else if (y_acc < rp.t_lo) accept = true;           // below every alpha
            else if (x == prev) { if (y_acc < rp.t_ret) accept = true; else draw = true; }
            else if (y_acc >= rp.t_hi) draw = true;            // above both remaining alphas
            else if (rp.t_in == rp.t_out) { if (y_acc < rp.t_in) accept = true; else draw = true; }   // q == 1 / popularity law: no test
            else {
                ++n_tests; n_probes += ceil_log2_p1(pdeg);
                key = ((uint64_t)(uint32_t)x << 32) | (uint32_t)prev;   // G.has_edge(x, prev) (:144)
                slot = mix64(key) & hash_mask;
                state = ST_PROBE;
            }
        } else if (state == ST_PROBE) {
            if (v == key) { if (y_acc < rp.t_in) accept = true; else draw = true; }
            else if (v == HASH_EMPTY) { if (y_acc < rp.t_out) accept = true; else draw = true; }
            else slot = (slot + 1) & hash_mask;
        }
        // ---- a new trial of step `len` (no memory access)
        if (draw) {
            const Philox4 r = philox4x32_10((uint32_t)wid, (uint32_t)(wid >> 32), (uint32_t)len, trial, k0, k1);
            ++trial;
            y_acc = r.z; y_al = r.y;
            kk = (int64_t)__umul64hi((uint64_t)r.x << 32, (uint64_t)K);   // floor(u*K)
            state = (WEIGHTED || (prev < 0 && first_slots != nullptr)) ? ST_ALIAS : ST_COL;
            if (rp.fold && prev >= 0) {
                // outlier: the return edge carries (1/p - B') extra area on top of the B'-high
                // dartboard of K unit-weight columns; re-drawn every trial, always accepted.
                const double u = (double)r.w * (1.0 / 4294967296.0);
                const double fm = WEIGHTED ? rp.fold_mass * w_ret : rp.fold_mass;
                const double area = rp.bound * (WEIGHTED ? w_row : (double)K) + fm;
                if (u * area < fm) { x = prev; accept = true; outlier = true; }
            }
            if (trial >= 100000u && prev >= 0) { x = prev; accept = true; outlier = true; }   // safety valve, unreachable for sane p, q
        }
        if (accept) {
            n_trials += trial;
            prev = cur; pdeg = K; cur = x;
            emit(cur);
            if (len >= L) finish();
            else if (wfold && !outlier) { slot = (uint64_t)(b + kk); state = ST_WRET; }   // weight of the arc just taken
            else state = ST_ROW;                  // (outlier: same edge walked back, w_ret unchanged)
        }
    }
    if (live) lens[i] = len;
    if (counters) {
        unsigned long long st = live && len > 0 ? (unsigned long long)(len - 1) : 0ull;
        unsigned long long c1 = n_trials, c2 = n_tests, c3 = n_probes;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            st += __shfl_xor_sync(0xFFFFFFFFu, st, o);
            c1 += __shfl_xor_sync(0xFFFFFFFFu, c1, o);
            c2 += __shfl_xor_sync(0xFFFFFFFFu, c2, o);
            c3 += __shfl_xor_sync(0xFFFFFFFFu, c3, o);
        }
        if (lane == 0) {
            atomicAdd(counters + 0, st);
            atomicAdd(counters + 1, c1);
            atomicAdd(counters + 2, c2);
            atomicAdd(counters + 3, c3);
        }
    }
}

}  // namespace n2v

using namespace n2v;

extern "C" int n2v_pack_rows(const int64_t *row_ptr, int32_t n_nodes, uint64_t *packed, int *overflow_flag,
                             void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    N2V_REQUIRE(n_nodes >= 0, "negative n_nodes");
    if (n_nodes == 0) return N2V_OK;
    N2V_REQUIRE(row_ptr && packed && overflow_flag, "NULL buffer");
    pack_rows_kernel<<<(n_nodes + 255) / 256, 256, 0, stream>>>(row_ptr, n_nodes, packed, overflow_flag);
    N2V_LAUNCH_CHECK();
    return N2V_OK;
}

extern "C" uint64_t n2v_edge_hash_capacity(int64_t nnz)
{
    uint64_t cap = 1024;
    while (cap < 2ull * (uint64_t)(nnz > 0 ? nnz : 1)) cap <<= 1;
    return cap;
}

extern "C" int n2v_edge_hash_build(const int64_t *row_ptr, const int32_t *col, int32_t n_nodes, int64_t nnz,
                                   unsigned long long *table, uint64_t capacity, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    N2V_REQUIRE(table && capacity >= 2 && (capacity & (capacity - 1)) == 0, "capacity must be a power of two");
    N2V_REQUIRE(capacity >= 2ull * (uint64_t)(nnz > 0 ? nnz : 1), "capacity below 2 * nnz");
    N2V_CHECK_CUDA(cudaMemsetAsync(table, 0xFF, sizeof(unsigned long long) * capacity, stream));
    if (nnz == 0) return N2V_OK;
    N2V_REQUIRE(row_ptr && col, "NULL buffer");
    edge_hash_build_kernel<<<(unsigned)((nnz + 255) / 256), 256, 0, stream>>>(row_ptr, col, n_nodes, nnz, table, capacity - 1);
    N2V_LAUNCH_CHECK();
    return N2V_OK;
}

extern "C" int n2v_walk_reject_indexed(const uint64_t *packed_rows, const int32_t *col, int64_t nnz, const double *w,
                                       const double *strength, const n2v_slot_t *node_slots, const unsigned long long *edge_hash,
                                       uint64_t hash_capacity, double p, double q, int symmetric,
                                       const int32_t *starts, int64_t n_walks, int32_t L, uint64_t seed,
                                       uint64_t walk_id_base, int32_t *walks, int32_t *lens,
                                       unsigned long long *counters, void *stream_)
{
    return n2v_walk_reject_indexed_law(packed_rows, col, nnz, w, strength, node_slots, nullptr, edge_hash, hash_capacity,
                                       p, q, symmetric, starts, n_walks, L, seed, walk_id_base, walks, lens, counters,
                                       stream_);
}

extern "C" int n2v_walk_reject_indexed_law(const uint64_t *packed_rows, const int32_t *col, int64_t nnz, const double *w,
                                           const double *strength, const n2v_slot_t *node_slots, const n2v_walk_law_t *law,
                                           const unsigned long long *edge_hash, uint64_t hash_capacity, double p, double q,
                                           int symmetric, const int32_t *starts, int64_t n_walks, int32_t L, uint64_t seed,
                                           uint64_t walk_id_base, int32_t *walks, int32_t *lens,
                                           unsigned long long *counters, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    N2V_REQUIRE(n_walks >= 0 && L >= 0, "negative size");
    N2V_REQUIRE(p > 0.0 && q > 0.0, "p and q must be positive");
    if (n_walks == 0 || L == 0) return N2V_OK;
    N2V_REQUIRE(packed_rows && col && edge_hash && starts && walks && lens, "NULL buffer");
    N2V_REQUIRE(hash_capacity >= 2 && (hash_capacity & (hash_capacity - 1)) == 0, "hash capacity must be a power of two");
    const bool pop_edges = law && law->pop_edges;
    const n2v_slot_t *first_slots = law ? law->first_slots : nullptr;
    N2V_REQUIRE(!(w || pop_edges) || node_slots, "weighted graph / popularity edge law needs node_slots");
    if (sm_count() <= 0) { set_error("no CUDA device"); return N2V_ECUDA; }
    // popularity edge law (node2vec.py:154-174): candidate ~ w / len(G[nbr]) from node_slots, alpha = 1/p on
    // the return edge and 1 elsewhere (q unused there); never folded (the row sums of w / pop are not supplied)
    const RejectParams rp = pop_edges ? make_reject_params(p, 1.0, true, symmetric, false)
                                      : make_reject_params(p, q, w != nullptr, symmetric, w != nullptr && strength != nullptr);
    const int64_t blocks = (n_walks + W2_BLOCK - 1) / W2_BLOCK;
    N2V_REQUIRE(blocks < 2147483647ll, "too many walks for one launch");
    if (w || pop_edges)
        walk_reject_indexed_kernel<true><<<(unsigned)blocks, W2_BLOCK, 0, stream>>>(
            packed_rows, col, node_slots, first_slots, w, strength, edge_hash, hash_capacity - 1, rp, nnz, starts, n_walks, L,
            (uint32_t)seed, (uint32_t)(seed >> 32), walk_id_base, walks, lens, counters);
    else
        walk_reject_indexed_kernel<false><<<(unsigned)blocks, W2_BLOCK, 0, stream>>>(
            packed_rows, col, node_slots, first_slots, w, strength, edge_hash, hash_capacity - 1, rp, nnz, starts, n_walks, L,
            (uint32_t)seed, (uint32_t)(seed >> 32), walk_id_base, walks, lens, counters);
    N2V_LAUNCH_CHECK();
    return N2V_OK;
}
