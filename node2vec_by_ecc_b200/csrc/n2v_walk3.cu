// Alias-mode walker over packed arc records.
//
// ncu on the first form (n2v_walk.cu: walk_alias_kernel, R-MAT 2^20, profiles/r01_i_*): DRAM 71 %
// busy, 7.6 sectors per step for 40 algorithmic bytes, 95 % long-scoreboard -- each step is 4-5
// random reads (row_ptr pair, etab_ptr, slot, col), three of them dependent. Here everything a
// step needs about the arc it just took sits in ONE 32-byte, sector-aligned record
//     { edge-table offset, row start << 24 | degree of the head, head node id }
// so a step is two dependent sector reads: the alias slot, then the record of the chosen arc.
// Same Philox addressing and float64 arithmetic as node2vec_walk/alias_draw (node2vec.py:55-79,
// :271-281): output is bit-identical to walk_alias_kernel.
#include "n2v_common.cuh"

namespace n2v {

struct __align__(32) ArcRec {
    long long tab_off;           // etab_ptr[e]
    unsigned long long row;      // row_ptr[col[e]] << 24 | deg(col[e])
    int32_t node;                // col[e]
    int32_t pad0; long long pad1;
};
static_assert(sizeof(ArcRec) == 32, "one sector per arc");
constexpr int W3_BLOCK = 256;
constexpr int W3_DEG_BITS = 24;

__global__ void pack_arcs_kernel(const int64_t *__restrict__ row_ptr, const int32_t *__restrict__ col,
                                 const int64_t *__restrict__ etab_ptr, int64_t nnz, ArcRec *__restrict__ recs,
                                 int *__restrict__ overflow)
{
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= nnz) return;
    const int32_t v = col[e];
    const int64_t b = row_ptr[v], d = row_ptr[v + 1] - b;
    if (d >= (1ll << W3_DEG_BITS) || b >= (1ll << (64 - W3_DEG_BITS))) *overflow = 1;
    ArcRec r;
    r.tab_off = etab_ptr[e];
    r.row = ((unsigned long long)b << W3_DEG_BITS) | (unsigned long long)(d & ((1ll << W3_DEG_BITS) - 1));
    r.node = v; r.pad0 = 0; r.pad1 = 0;
    recs[e] = r;
}

__global__ void __launch_bounds__(W3_BLOCK)
walk_alias_packed_kernel(const unsigned long long *__restrict__ packed_rows, const n2v_slot_t *__restrict__ node_slots,
                         const ArcRec *__restrict__ recs, const n2v_slot_t *__restrict__ edge_slots,
                         const int32_t *__restrict__ starts, int64_t n_walks, int32_t L, uint32_t k0, uint32_t k1,
                         uint64_t walk_id_base, int32_t *__restrict__ walks, int32_t *__restrict__ lens)
{
    __shared__ int32_t stage_all[W3_BLOCK][8 + 1];
    int32_t *stage = stage_all[threadIdx.x];
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_walks) return;
    const uint64_t wid = walk_id_base + (uint64_t)i;
    int32_t *const out = walks + i * (int64_t)L;
    const int32_t start = starts[i];
    int32_t len = 1;
    stage[0] = start;
    // first step: node table of the start node (node2vec.py:69-70)
    unsigned long long row = __ldg(packed_rows + start);
    const n2v_slot_t *tab = node_slots + (int64_t)(row >> W3_DEG_BITS);
    for (int32_t s = 1; s < L; ++s) {
        const int64_t b = (int64_t)(row >> W3_DEG_BITS);
        const int64_t K = (int64_t)(row & ((1ull << W3_DEG_BITS) - 1));
        if (K <= 0) break;                                  // dead end (:76-77)
        const Philox4 r = philox4x32_10((uint32_t)wid, (uint32_t)(wid >> 32), (uint32_t)s, 0u, k0, k1);
        // kk = int(np.floor(np.random.rand()*K)) (:277), u1 = r.x * 2^-32 exactly
        const int64_t kk = (int64_t)floor(__dmul_rn((double)r.x * (1.0 / 4294967296.0), (double)K));
        const uint2 sl = __ldg(reinterpret_cast<const uint2 *>(tab + kk));
        const int64_t k = (r.y < sl.y) ? kk : (int64_t)(int32_t)sl.x;          // :278-281
        const ArcRec *rp = recs + (b + k);
        const longlong2 lo = __ldg(reinterpret_cast<const longlong2 *>(rp));    // {tab_off, row}
        const int32_t node = __ldg(&rp->node);
        tab = edge_slots + lo.x;
        row = (unsigned long long)lo.y;
        stage[len & 7] = node;
        ++len;
        if ((len & 7) == 0) {
#pragma unroll
            for (int c = 0; c < 8; ++c) out[len - 8 + c] = stage[c];
        }
    }
    for (int32_t t = len & ~7; t < L; ++t) out[t] = t < len ? stage[t & 7] : -1;
    lens[i] = L > 0 ? len : 0;
}

}  // namespace n2v

using namespace n2v;

extern "C" size_t n2v_arc_record_bytes(void) { return sizeof(ArcRec); }

extern "C" int n2v_pack_arcs(const int64_t *row_ptr, const int32_t *col, const int64_t *etab_ptr, int64_t nnz,
                             void *recs, int *overflow_flag, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    N2V_REQUIRE(nnz >= 0, "negative nnz");
    if (nnz == 0) return N2V_OK;
    N2V_REQUIRE(row_ptr && col && etab_ptr && recs && overflow_flag, "NULL buffer");
    N2V_REQUIRE(((uintptr_t)recs & 31) == 0, "recs must be 32-byte aligned");
    pack_arcs_kernel<<<(unsigned)((nnz + 255) / 256), 256, 0, stream>>>(row_ptr, col, etab_ptr, nnz, (ArcRec *)recs,
                                                                        overflow_flag);
    N2V_LAUNCH_CHECK();
    return N2V_OK;
}

extern "C" int n2v_walk_alias_packed(const uint64_t *packed_rows, const n2v_slot_t *node_slots, const void *recs,
                                     const n2v_slot_t *edge_slots, const int32_t *starts, int64_t n_walks,
                                     int32_t L, uint64_t seed, uint64_t walk_id_base, int32_t *walks,
                                     int32_t *lens, void *stream_)
{
    cudaStream_t stream = (cudaStream_t)stream_;
    N2V_REQUIRE(n_walks >= 0 && L >= 0, "negative size");
    if (n_walks == 0 || L == 0) return N2V_OK;
    N2V_REQUIRE(packed_rows && node_slots && starts && walks && lens, "NULL buffer");
    N2V_REQUIRE(L <= 1 || recs, "arc records are NULL");                  // read from the first step on
    N2V_REQUIRE(L <= 2 || edge_slots, "edge tables are NULL");              // first used at the second step
    if (sm_count() <= 0) { set_error("no CUDA device"); return N2V_ECUDA; }
    const int64_t blocks = (n_walks + W3_BLOCK - 1) / W3_BLOCK;
    N2V_REQUIRE(blocks < 2147483647ll, "too many walks for one launch");
    walk_alias_packed_kernel<<<(unsigned)blocks, W3_BLOCK, 0, stream>>>(
        (const unsigned long long *)packed_rows, node_slots, (const ArcRec *)recs, edge_slots, starts, n_walks, L,
        (uint32_t)seed, (uint32_t)(seed >> 32), walk_id_base, walks, lens);
    N2V_LAUNCH_CHECK();
    return N2V_OK;
}
