// Alias-table builder (north-star subsystem 2): exact restatement of the reference's
// alias_setup (src/node2vec.py:240-269) applied to every node (:184-188) and every arc
// (get_alias_edge, :133-152), one warp per table.
//
// Exactness. J depends on the LIFO pairing order of the two Python lists and q on the float64
// evaluation order, so the pairing loop cannot be parallelised inside a table. What CAN be said
// about the two stacks: `smaller` is always (a prefix of the initial ascending small list) plus
// at most one demoted large on top; `larger` is a prefix of the initial ascending large list
// whose top may carry a modified q. So no stack storage is needed: two downward scans over class
// marks + one "pending" register reproduce the pop order exactly. Phases per table:
//   A  lanes: un[k] = w (/p | /q by the distance-1 test), strided, coalesced on the row
//   B  all lanes redundantly: norm = sequential left-to-right float64 sum (Python 2 sum())
//   C  lanes: q[k] = K * (un[k] / norm); mark small (-1) / large (-2)
//   D  lane 0: the pairing loop (two pointers + pending)
//   E  lanes: pack {alias, ceil(q*2^32)} slots
// Tables of <= 32 entries run B..E entirely in registers (finish_table_small).
// All float64 operations use the _rn intrinsics: no FMA contraction, same bits as CPython.
// HBM-bound roofline: 8 B/entry slot write + 12 B/entry scratch traffic (L2-resident per chunk).
#include <cub/cub.cuh>

#include "n2v_common.cuh"

namespace n2v {

constexpr int MARK_SMALL = -1;
constexpr int MARK_LARGE = -2;
constexpr int MARK_DEMOTED = -3;

// which arc does CSR position e belong to: largest u with row_ptr[u] <= e
__device__ __forceinline__ int32_t row_of_arc(const int64_t *__restrict__ row_ptr, int32_t n, int64_t e)
{
    int32_t lo = 0, hi = n;   // invariant: row_ptr[lo] <= e < row_ptr[hi]
    while (hi - lo > 1) {
        int32_t mid = lo + ((hi - lo) >> 1);
        if (__ldg(row_ptr + mid) <= e) lo = mid; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ double shfl_f64(double v, int src)
{
    return __hiloint2double(__shfl_sync(0xFFFFFFFFu, __double2hiint(v), src), __shfl_sync(0xFFFFFFFFu, __double2loint(v), src));
}

// Phases B..E for a table of at most 32 entries, entirely in registers: lane k owns entry k (its q and
// its J), the two stacks are two ballots, the pairing loop runs warp-uniformly on shuffled values. Same
// operations in the same order as finish_table below (and as node2vec.py:240-269), no memory traffic
// inside the loop. Most tables of a sparse graph are this small (C2: mean degree 33 after the split).
__device__ __forceinline__ void finish_table_small(int K, double *wq, int32_t *wJ, n2v_slot_t *slots, int lane,
                                                   bool normalize)
{
    __syncwarp();
    const bool mine = lane < K;
    double qk = mine ? wq[lane] : 0.0;
    double norm = 0.0;
    if (normalize)
        for (int j = 0; j < K; ++j) norm = __dadd_rn(norm, shfl_f64(qk, j));            // B: left to right
    qk = __dmul_rn((double)K, normalize ? __ddiv_rn(qk, norm) : qk);                    // C
    uint32_t smask = __ballot_sync(0xFFFFFFFFu, mine && qk < 1.0);
    uint32_t lmask = __ballot_sync(0xFFFFFFFFu, mine && !(qk < 1.0));
    int32_t Jk = 0;                                                                     // leftovers keep J = 0 (:248)
    int pending = -1, cur_large = -1;
    double q_large = 0.0, q_pending = 0.0;
    for (;;) {                                                                          // D (:259-268)
        int small; double q_small;
        if (pending >= 0) { small = pending; q_small = q_pending; }
        else {
            if (smask == 0) break;
            small = 31 - __clz(smask);                      // top of `smaller`: highest index first
            q_small = shfl_f64(qk, small);
        }
        if (cur_large < 0) {
            if (lmask == 0) break;
            cur_large = 31 - __clz(lmask); lmask &= ~(1u << cur_large);
            q_large = shfl_f64(qk, cur_large);
        }
        if (pending >= 0) pending = -1; else smask &= ~(1u << small);
        if (lane == small) Jk = cur_large;                                              // J[small] = large
        q_large = __dadd_rn(__dadd_rn(q_large, q_small), -1.0);                         // q[large]+q[small]-1.0
        if (q_large < 1.0) {
            if (lane == cur_large) qk = q_large;
            pending = cur_large; q_pending = q_large; cur_large = -1;
        }
    }
    if (cur_large >= 0 && lane == cur_large) qk = q_large;
    if (mine) {                                                                         // E
        wq[lane] = qk; wJ[lane] = Jk;
        slots[lane] = make_slot(lane, Jk, qk);
    }
    __syncwarp();
}

// Phases B..E on a table whose un-normalised probabilities are already in wq[0..K)
__device__ __forceinline__ void finish_table(int64_t K, double *wq, int32_t *wJ, n2v_slot_t *slots, int lane,
                                             bool normalize = true)
{
    if (K <= 32) { finish_table_small((int)K, wq, wJ, slots, lane, normalize); return; }
    __syncwarp();
    // B: norm_const = sum(unnormalized_probs), left to right (node2vec.py:148,:186). 32 values per
    // coalesced load; every lane then adds them in index order (identical norm on all lanes).
    double norm = 0.0;
    if (normalize) {
        for (int64_t k0 = 0; k0 < K; k0 += 32) {
            const double v = (k0 + lane < K) ? wq[k0 + lane] : 0.0;
            const int m = (K - k0 < 32) ? (int)(K - k0) : 32;
#pragma unroll 8
            for (int j = 0; j < m; ++j) norm = __dadd_rn(norm, __shfl_sync(0xFFFFFFFFu, v, j));
        }
    }
    // C: normalized = float(u)/norm (:149); q[kk] = K*prob (:253); classify (:254-257)
    const double Kd = (double)K;
    for (int64_t k = lane; k < K; k += 32) {
        double qq = __dmul_rn(Kd, normalize ? __ddiv_rn(wq[k], norm) : wq[k]);
        wq[k] = qq;
        wJ[k] = (qq < 1.0) ? MARK_SMALL : MARK_LARGE;
    }
    // D: while len(smaller) > 0 and len(larger) > 0 (:259-268). Warp-uniform: every lane runs the
    // same control flow on the same registers; the two downward scans read the class marks 32 at a
    // time (one coalesced load + ballot per block) and pop them with clz; q of the running large
    // and of a demoted large stay in registers; lane 0 does the stores.
    {
        auto block_mask = [&](int64_t blk, int mark) -> uint32_t {
            __syncwarp();                                   // lane 0's stores visible to the block load
            const int64_t k = (blk << 5) + lane;
            return __ballot_sync(0xFFFFFFFFu, k < K && wJ[k] == mark);
        };
        int64_t sb = (K - 1) >> 5, lb = sb;
        uint32_t smask = block_mask(sb, MARK_SMALL), lmask = block_mask(lb, MARK_LARGE);
        int64_t pending = -1, cur_large = -1;
        double q_large = 0.0, q_pending = 0.0;
        for (;;) {
            int64_t small; double q_small;
            if (pending >= 0) { small = pending; q_small = q_pending; }
            else {
                while (smask == 0 && sb > 0) { --sb; smask = block_mask(sb, MARK_SMALL); }
                if (smask == 0) break;
                small = (sb << 5) + (31 - __clz(smask));    // top of `smaller`: highest index first
                q_small = wq[small];
            }
            if (cur_large < 0) {
                while (lmask == 0 && lb > 0) { --lb; lmask = block_mask(lb, MARK_LARGE); }
                if (lmask == 0) break;
                const int bit = 31 - __clz(lmask);
                cur_large = (lb << 5) + bit; lmask &= ~(1u << bit);
                q_large = wq[cur_large];
            }
            if (pending >= 0) pending = -1; else smask &= ~(1u << (int)(small & 31));
            if (lane == 0) wJ[small] = (int32_t)cur_large;                   // J[small] = large
            q_large = __dadd_rn(__dadd_rn(q_large, q_small), -1.0);          // q[large]+q[small]-1.0
            if (q_large < 1.0) {
                if (lane == 0) { wq[cur_large] = q_large; wJ[cur_large] = MARK_DEMOTED; }
                pending = cur_large; q_pending = q_large; cur_large = -1;
            }
        }
        if (cur_large >= 0 && lane == 0) wq[cur_large] = q_large;
    }
    __syncwarp();
    // E: leftovers keep J = 0 (np.zeros, :248); pack
    for (int64_t k = lane; k < K; k += 32) {
        int32_t J = wJ[k];
        if (J < 0) { J = 0; wJ[k] = 0; }
        slots[k] = make_slot((int32_t)k, J, wq[k]);
    }
}

// ---- node tables -----------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
alias_nodes_kernel(const int64_t *__restrict__ row_ptr, const int32_t *__restrict__ col,
                   const double *__restrict__ w, int32_t n_nodes, const uint8_t *__restrict__ is_item,
                   int popwalk, n2v_slot_t *__restrict__ slots, int32_t *__restrict__ work_J,
                   double *__restrict__ work_q)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t v = warp; v < n_nodes; v += n_warps) {
        const int64_t b = row_ptr[v], K = row_ptr[v + 1] - b;
        if (K <= 0) continue;
        const bool pop = (popwalk & 1) && !(is_item && is_item[v]);
        for (int64_t k = lane; k < K; k += 32) {
            double u = w ? w[b + k] : 1.0;
            if (pop) {   // weight*1.0/len(G[nbr]) (:21,:218)
                int32_t nb = col[b + k];
                u = __ddiv_rn(__dmul_rn(u, 1.0), (double)(row_ptr[nb + 1] - row_ptr[nb]));
            }
            work_q[b + k] = u;
        }
        finish_table(K, work_q + b, work_J + b, slots + b, lane, !(popwalk & 2));
    }
}

// ---- edge tables -----------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
alias_edges_kernel(const int64_t *__restrict__ row_ptr, const int32_t *__restrict__ col,
                   const double *__restrict__ w, int32_t n_nodes, double p, double q, int symmetric,
                   int popwalk, const int64_t *__restrict__ etab_ptr, int64_t arc_begin, int64_t arc_end,
                   n2v_slot_t *__restrict__ slots, int32_t *__restrict__ work_J,
                   double *__restrict__ work_q)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t base = etab_ptr[arc_begin];
    for (int64_t e = arc_begin + warp; e < arc_end; e += n_warps) {
        const int32_t src = row_of_arc(row_ptr, n_nodes, e);
        const int32_t dst = col[e];
        const int64_t b = row_ptr[dst], K = row_ptr[dst + 1] - b;
        if (K <= 0) continue;
        const int64_t off = etab_ptr[e];
        double *wq = work_q + (off - base);
        int32_t *wJ = work_J + (off - base);
        const int64_t sb = row_ptr[src], se = row_ptr[src + 1];
        for (int64_t k = lane; k < K; k += 32) {
            const int32_t nbr = col[b + k];
            const double wt = w ? w[b + k] : 1.0;
            double u;
            if (popwalk) {
                // get_alias_edge_pop (:154-174): pop = len(G[dst_nbr]); w/(p*pop) for the return
                // edge, w/pop otherwise (both other branches, q does not appear)
                const double pop = (double)(row_ptr[nbr + 1] - row_ptr[nbr]);
                u = (nbr == src) ? __ddiv_rn(wt, __dmul_rn(p, pop)) : __ddiv_rn(wt, pop);
            } else if (nbr == src) u = __ddiv_rn(wt, p);                           // :142-143
            else {
                // G.has_edge(dst_nbr, src) (:144): arc nbr->src. On a symmetric CSR that is
                // nbr in adj(src): one row for the whole table, cache friendly.
                bool d1 = symmetric ? sorted_contains(col, sb, se, nbr)
                                    : sorted_contains(col, row_ptr[nbr], row_ptr[nbr + 1], src);
                u = d1 ? wt : __ddiv_rn(wt, q);                                    // :145-147
            }
            wq[k] = u;
        }
        finish_table(K, wq, wJ, slots + off, lane);
    }
}

__global__ void etab_sizes_kernel(const int64_t *__restrict__ row_ptr, const int32_t *__restrict__ col,
                                  int64_t nnz, int64_t *__restrict__ sizes)
{
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e > nnz) return;
    if (e == nnz) { sizes[e] = 0; return; }
    int32_t v = col[e];
    sizes[e] = row_ptr[v + 1] - row_ptr[v];
}

}  // namespace n2v

using namespace n2v;

extern "C" size_t n2v_etab_workspace_bytes(int64_t nnz)
{
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, (int64_t *)nullptr, (int64_t *)nullptr, nnz + 1);
    return tb + 256;
}

extern "C" int n2v_etab_offsets(const int64_t *row_ptr, const int32_t *col, int32_t n_nodes,
                                int64_t nnz, int64_t *etab_ptr, void *workspace,
                                size_t workspace_bytes, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    (void)n_nodes;
    N2V_REQUIRE(row_ptr && etab_ptr && workspace && nnz >= 0, "bad argument");
    N2V_REQUIRE(nnz == 0 || col, "col is NULL");
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, (int64_t *)nullptr, (int64_t *)nullptr, nnz + 1);
    if (tb > workspace_bytes) { set_error("etab workspace too small: need %zu", tb); return N2V_ENOMEM; }
    const int T = 256;
    etab_sizes_kernel<<<(unsigned)((nnz + 1 + T - 1) / T), T, 0, stream>>>(row_ptr, col, nnz, etab_ptr);
    N2V_LAUNCH_CHECK();
    N2V_CHECK_CUDA(cub::DeviceScan::ExclusiveSum(workspace, tb, etab_ptr, etab_ptr, nnz + 1, stream));
    return N2V_OK;
}

static int table_grid(int64_t n_tables)
{
    int sms = sm_count();
    if (sms <= 0) return -1;
    int64_t blocks = (n_tables + 7) / 8;          // 8 warps per block
    int64_t cap = (int64_t)sms * 8;               // 8 blocks x 8 warps = 64 warps / SM
    return (int)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

extern "C" int n2v_alias_build_nodes(const int64_t *row_ptr, const int32_t *col, const double *w,
                                     int32_t n_nodes, const uint8_t *is_item, int popwalk,
                                     n2v_slot_t *slots, int32_t *work_J, double *work_q,
                                     void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    N2V_REQUIRE(n_nodes >= 0, "negative n_nodes");
    if (n_nodes == 0) return N2V_OK;
    N2V_REQUIRE(row_ptr && col && slots && work_J && work_q, "NULL buffer");
    int grid = table_grid(n_nodes);
    if (grid < 0) { set_error("no CUDA device"); return N2V_ECUDA; }
    alias_nodes_kernel<<<grid, 256, 0, stream>>>(row_ptr, col, w, n_nodes, is_item, popwalk, slots,
                                                 work_J, work_q);
    N2V_LAUNCH_CHECK();
    return N2V_OK;
}

extern "C" int n2v_alias_build_edges(const int64_t *row_ptr, const int32_t *col, const double *w,
                                     int32_t n_nodes, double p, double q, int symmetric, int popwalk,
                                     const int64_t *etab_ptr, int64_t arc_begin, int64_t arc_end,
                                     n2v_slot_t *slots, int32_t *work_J, double *work_q,
                                     void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    N2V_REQUIRE(arc_begin >= 0 && arc_end >= arc_begin, "bad arc range");
    N2V_REQUIRE(p > 0.0 && q > 0.0, "p and q must be positive");
    if (arc_end == arc_begin) return N2V_OK;
    N2V_REQUIRE(row_ptr && col && etab_ptr && slots && work_J && work_q, "NULL buffer");
    int grid = table_grid(arc_end - arc_begin);
    if (grid < 0) { set_error("no CUDA device"); return N2V_ECUDA; }
    alias_edges_kernel<<<grid, 256, 0, stream>>>(row_ptr, col, w, n_nodes, p, q, symmetric, popwalk, etab_ptr,
                                                 arc_begin, arc_end, slots, work_J, work_q);
    N2V_LAUNCH_CHECK();
    return N2V_OK;
}
