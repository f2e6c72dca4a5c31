// Second-order biased random walks (north-star subsystem 3).
//
// Work decomposition: one lane per walker, a warp owns 32 consecutive walk ids. A walk step is a
// chain of 2-3 dependent random HBM reads and nothing else, so throughput is set by the number of
// independent chains in flight per SM (2048 here), not by lanes per walker; a warp-per-walker
// layout would leave 31/32 of that memory-level parallelism unused. What IS done per warp is the
// output: tokens are staged in shared memory and flushed as full 32-byte sectors (8 steps of one
// walker per sector), so the [n_walks, L] walk-major corpus the SGNS kernel wants is written with
// no partial-sector traffic.
//
// alias mode    : bit-exact restatement of node2vec_walk (src/node2vec.py:55-79) + alias_draw
//                 (:271-281); per step reads row_ptr pair, etab_ptr, one 8-byte slot, one col id.
// rejection mode: KnightKing-style dartboard on the same transition law (get_alias_edge,
//                 :142-150) with pre-accept / pre-reject bounds and the return edge folded out as
//                 an outlier; needs no edge tables (graphs whose Sigma deg^2 does not fit).
#include "n2v_common.cuh"
#include "n2v_reject.cuh"

namespace n2v {

constexpr int WALK_BLOCK = 256;
constexpr int STAGE = 8;   // tokens per walker per flush = one 32-byte sector

// Flush the warp's staged tokens: stage[lane][0..STAGE) -> walks[(w0+lane)*L + s0 + 0..STAGE).
// Lanes are re-mapped so that 8 consecutive lanes write one walker's 32-byte sector.
__device__ __forceinline__ void flush_stage(int32_t (*stage)[STAGE + 1], int32_t *__restrict__ walks,
                                            int64_t w0, int64_t n_walks, int32_t L, int32_t s0, int lane)
{
    __syncwarp();
    const int c = lane & 7;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int wl = (lane >> 3) + 4 * r;            // walker within the warp (4 sectors per pass)
        const int64_t wi = w0 + wl;
        const int32_t s = s0 + c;
        if (wi < n_walks && s < L) walks[wi * L + s] = stage[wl][c];
    }
    __syncwarp();
}

__global__ void __launch_bounds__(WALK_BLOCK)
walk_alias_kernel(const int64_t *__restrict__ row_ptr, const int32_t *__restrict__ col,
                  const n2v_slot_t *__restrict__ node_slots, const int64_t *__restrict__ etab_ptr,
                  const n2v_slot_t *__restrict__ edge_slots, const int32_t *__restrict__ starts,
                  int64_t n_walks, int32_t L, uint32_t k0, uint32_t k1, uint64_t walk_id_base,
                  int32_t *__restrict__ walks, int32_t *__restrict__ lens)
{
    __shared__ int32_t stage_all[WALK_BLOCK / 32][32][STAGE + 1];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    int32_t (*stage)[STAGE + 1] = stage_all[wib];
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t w0 = i - lane;
    const bool live = i < n_walks;
    const uint64_t wid = walk_id_base + (uint64_t)i;

    int32_t cur = live ? starts[i] : -1;
    int64_t arc = -1;
    int32_t len = live ? 1 : 0;
    bool alive = live;
    stage[lane][0] = cur;
    for (int32_t s = 1; s < L; ++s) {                      // s = index of the token being drawn
        int32_t tok = -1;
        if (alive) {
            const int64_t b = __ldg(row_ptr + cur), K = __ldg(row_ptr + cur + 1) - b;
            if (K <= 0) alive = false;                     // dead end: break (:76-77)
            else {
                const Philox4 r = philox4x32_10((uint32_t)wid, (uint32_t)(wid >> 32), (uint32_t)s, 0u, k0, k1);
                // kk = int(np.floor(np.random.rand()*K)) (:277), u1 = r.x * 2^-32 exactly
                const int64_t kk = (int64_t)floor(__dmul_rn((double)r.x * (1.0 / 4294967296.0), (double)K));
                const n2v_slot_t *tab = (s == 1) ? node_slots + b : edge_slots + __ldg(etab_ptr + arc);
                const uint2 sl = __ldg(reinterpret_cast<const uint2 *>(tab + kk));
                const int64_t k = (r.y < sl.y) ? kk : (int64_t)(int32_t)sl.x;   // :278-281
                arc = b + k;
                cur = __ldg(col + arc);
                tok = cur;
                ++len;
            }
        }
        stage[lane][s & (STAGE - 1)] = tok;
        if ((s & (STAGE - 1)) == STAGE - 1) flush_stage(stage, walks, w0, n_walks, L, s - (STAGE - 1), lane);
    }
    if (L & (STAGE - 1)) flush_stage(stage, walks, w0, n_walks, L, L & ~(STAGE - 1), lane);
    if (live) lens[i] = (L > 0) ? len : 0;
}

// ---- rejection mode ----------------------------------------------------------------------------
template <bool WEIGHTED>
__global__ void __launch_bounds__(WALK_BLOCK)
walk_reject_kernel(const int64_t *__restrict__ row_ptr, const int32_t *__restrict__ col,
                   const n2v_slot_t *__restrict__ node_slots, const n2v_slot_t *__restrict__ first_slots,
                   RejectParams rp, int symmetric,
                   const int32_t *__restrict__ starts, int64_t n_walks, int32_t L, uint32_t k0,
                   uint32_t k1, uint64_t walk_id_base, int32_t *__restrict__ walks,
                   int32_t *__restrict__ lens, unsigned long long *__restrict__ counters)
{
    __shared__ int32_t stage_all[WALK_BLOCK / 32][32][STAGE + 1];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    int32_t (*stage)[STAGE + 1] = stage_all[wib];
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t w0 = i - lane;
    const bool live = i < n_walks;
    const uint64_t wid = walk_id_base + (uint64_t)i;

    int32_t cur = live ? starts[i] : -1, prev = -1;
    int64_t pb = 0, pe = 0;                                // row of prev
    int32_t len = live ? 1 : 0;
    bool alive = live;
    unsigned long long n_trials = 0, n_tests = 0, n_probes = 0;
    stage[lane][0] = cur;
    for (int32_t s = 1; s < L; ++s) {
        int32_t tok = -1;
        if (alive) {
            const int64_t b = __ldg(row_ptr + cur), K = __ldg(row_ptr + cur + 1) - b;
            if (K <= 0) alive = false;
            else {
                int32_t nxt = -1;
                uint32_t trial = 0;
                while (nxt < 0) {
                    const Philox4 r = philox4x32_10((uint32_t)wid, (uint32_t)(wid >> 32), (uint32_t)s, trial, k0, k1);
                    ++trial;
                    if (rp.fold && prev >= 0) {
                        // outlier: the return edge carries (1/p - B') extra area on top of the
                        // B'-high dartboard of K unit-weight columns; a dart lands there with
                        // that share of the total area and is always accepted (re-drawn per trial).
                        const double u = (double)r.w * (1.0 / 4294967296.0);
                        if (u * (rp.bound * (double)K + rp.fold_mass) < rp.fold_mass) { nxt = prev; break; }
                    }
                    int64_t k = (int64_t)__umul64hi((uint64_t)r.x << 32, (uint64_t)K);   // floor(u*K)
                    // static law: the node alias table ~ w(cur, .) (weighted graphs / the popularity
                    // edge law); the first step may have its own table (popularity node tables, :213-218)
                    const n2v_slot_t *tab = (prev < 0 && first_slots) ? first_slots : (WEIGHTED ? node_slots : nullptr);
                    if (tab) {
                        const uint2 sl = __ldg(reinterpret_cast<const uint2 *>(tab + b + k));
                        if (!(r.y < sl.y)) k = (int64_t)(int32_t)sl.x;
                    }
                    const int32_t x = __ldg(col + b + k);
                    if (prev < 0) { nxt = x; break; }          // first step: one static draw (:69-70)
                    const uint32_t y = r.z;
                    if (y < rp.t_lo) { nxt = x; break; }       // pre-accept: below every alpha
                    if (x == prev) { if (y < rp.t_ret) nxt = x; }
                    else if (y >= rp.t_hi) { /* pre-reject: above both remaining alphas */ }
                    else if (rp.t_in == rp.t_out) { if (y < rp.t_in) nxt = x; }   // q == 1 / popularity law: no test
                    else {
                        ++n_tests;
                        n_probes += ceil_log2_p1(pe - pb);
                        // distance-1 test, G.has_edge(x, prev) (:144)
                        const bool d1 = symmetric ? sorted_contains(col, pb, pe, x)
                                                  : sorted_contains(col, __ldg(row_ptr + x), __ldg(row_ptr + x + 1), prev);
                        if (y < (d1 ? rp.t_in : rp.t_out)) nxt = x;
                    }
                    if (trial >= 100000u && nxt < 0) nxt = x;   // safety valve, never reached for sane p, q
                }
                n_trials += trial;
                prev = cur; pb = b; pe = b + K;
                cur = nxt;
                tok = cur;
                ++len;
            }
        }
        stage[lane][s & (STAGE - 1)] = tok;
        if ((s & (STAGE - 1)) == STAGE - 1) flush_stage(stage, walks, w0, n_walks, L, s - (STAGE - 1), lane);
    }
    if (L & (STAGE - 1)) flush_stage(stage, walks, w0, n_walks, L, L & ~(STAGE - 1), lane);
    if (live) lens[i] = (L > 0) ? len : 0;
    if (counters) {
        unsigned long long st = live && len > 0 ? (unsigned long long)(len - 1) : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            st += __shfl_xor_sync(0xFFFFFFFFu, st, o);
            n_trials += __shfl_xor_sync(0xFFFFFFFFu, n_trials, o);
            n_tests += __shfl_xor_sync(0xFFFFFFFFu, n_tests, o);
            n_probes += __shfl_xor_sync(0xFFFFFFFFu, n_probes, o);
        }
        if (lane == 0) {
            atomicAdd(counters + 0, st);
            atomicAdd(counters + 1, n_trials);
            atomicAdd(counters + 2, n_tests);
            atomicAdd(counters + 3, n_probes);
        }
    }
}

}  // namespace n2v

using namespace n2v;

extern "C" int n2v_walk_alias(const int64_t *row_ptr, const int32_t *col, const n2v_slot_t *node_slots,
                              const int64_t *etab_ptr, const n2v_slot_t *edge_slots,
                              const int32_t *starts, int64_t n_walks, int32_t L, uint64_t seed,
                              uint64_t walk_id_base, int32_t *walks, int32_t *lens, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    N2V_REQUIRE(n_walks >= 0 && L >= 0, "negative size");
    if (n_walks == 0 || L == 0) return N2V_OK;
    N2V_REQUIRE(row_ptr && col && node_slots && starts && walks && lens, "NULL buffer");
    N2V_REQUIRE(L <= 2 || (etab_ptr && edge_slots), "edge tables are NULL");
    if (sm_count() <= 0) { set_error("no CUDA device"); return N2V_ECUDA; }
    const int64_t blocks = (n_walks + WALK_BLOCK - 1) / WALK_BLOCK;
    N2V_REQUIRE(blocks < 2147483647ll, "too many walks for one launch");
    walk_alias_kernel<<<(unsigned)blocks, WALK_BLOCK, 0, stream>>>(
        row_ptr, col, node_slots, etab_ptr, edge_slots, starts, n_walks, L, (uint32_t)seed,
        (uint32_t)(seed >> 32), walk_id_base, walks, lens);
    N2V_LAUNCH_CHECK();
    return N2V_OK;
}

extern "C" int n2v_walk_reject(const int64_t *row_ptr, const int32_t *col, const double *w,
                               const n2v_slot_t *node_slots, double p, double q, int symmetric,
                               const int32_t *starts, int64_t n_walks, int32_t L, uint64_t seed,
                               uint64_t walk_id_base, int32_t *walks, int32_t *lens,
                               unsigned long long *counters, void *stream_)
{
    return n2v_walk_reject_law(row_ptr, col, w, node_slots, nullptr, p, q, symmetric, starts, n_walks, L, seed,
                               walk_id_base, walks, lens, counters, stream_);
}

extern "C" int n2v_walk_reject_law(const int64_t *row_ptr, const int32_t *col, const double *w,
                                   const n2v_slot_t *node_slots, const n2v_walk_law_t *law, double p, double q,
                                   int symmetric, const int32_t *starts, int64_t n_walks, int32_t L, uint64_t seed,
                                   uint64_t walk_id_base, int32_t *walks, int32_t *lens,
                                   unsigned long long *counters, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    N2V_REQUIRE(n_walks >= 0 && L >= 0, "negative size");
    N2V_REQUIRE(p > 0.0 && q > 0.0, "p and q must be positive");
    if (n_walks == 0 || L == 0) return N2V_OK;
    N2V_REQUIRE(row_ptr && col && starts && walks && lens, "NULL buffer");
    const bool pop_edges = law && law->pop_edges;
    const n2v_slot_t *first_slots = law ? law->first_slots : nullptr;
    N2V_REQUIRE(!(w || pop_edges) || node_slots, "weighted graph / popularity edge law needs node_slots");
    if (sm_count() <= 0) { set_error("no CUDA device"); return N2V_ECUDA; }
    // popularity edge law (node2vec.py:154-174): candidate ~ w / len(G[nbr]) from node_slots, alpha = 1/p on
    // the return edge and 1 elsewhere (q is not used by the reference there); no folding (row sums unknown)
    const RejectParams rp = pop_edges ? make_reject_params(p, 1.0, true, symmetric)
                                      : make_reject_params(p, q, w != nullptr, symmetric);
    const int64_t blocks = (n_walks + WALK_BLOCK - 1) / WALK_BLOCK;
    N2V_REQUIRE(blocks < 2147483647ll, "too many walks for one launch");
    if (w || pop_edges)
        walk_reject_kernel<true><<<(unsigned)blocks, WALK_BLOCK, 0, stream>>>(
            row_ptr, col, node_slots, first_slots, rp, symmetric, starts, n_walks, L, (uint32_t)seed,
            (uint32_t)(seed >> 32), walk_id_base, walks, lens, counters);
    else
        walk_reject_kernel<false><<<(unsigned)blocks, WALK_BLOCK, 0, stream>>>(
            row_ptr, col, node_slots, first_slots, rp, symmetric, starts, n_walks, L, (uint32_t)seed,
            (uint32_t)(seed >> 32), walk_id_base, walks, lens, counters);
    N2V_LAUNCH_CHECK();
    return N2V_OK;
}
