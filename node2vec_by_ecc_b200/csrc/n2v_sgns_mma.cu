// EXPERIMENT (north-star subsystem 4: "an ncu-evidenced choice against a small batched-GEMM variant"):
// the shared-negative law of n2v_sgns.cu's v3 kernel with the arithmetic of one centre's window done as
// three small GEMMs on the tensor cores (mma.sync m16n8k8 TF32, fp32 accumulate):
//     C = syn0 rows of the centre's contexts  [n <= 32 x 128]      T = syn1neg rows of the 6 targets [8 x 128]
//     F  = C T^t                [n x 8]     (16 k-steps)
//     G  = (label - sigma(F)) * alpha, masked
//     dT = G^t C                [8 x 128]   as dT^t = C^t G
//     dC = G T                  [n x 128]
// then one red.global.add.v4.f32 per context row and per target row -- the same rows v3 moves.
// What it gives up, and why it is not the default: (i) window-batch semantics -- every dot
// of the window is taken before any update (pWord2Vec's scheme), not gensim's pair-by-pair order, so a
// one-warp run no longer equals the oracle; (ii) TF32 rounds the operands to 10 mantissa bits (3xTF32
// would triple the MMA count); (iii) both operand tiles live in shared memory: 22 KB per warp, 8 warps
// per SM instead of 20. Selected by n2v_sgns_params_t.tuning bit 4 (N2V_SGNS_TUNING=16); timed by
// bench.py --sgns-variant mma; profiles/r02_*_mma*.
#include "n2v_common.cuh"
#include "n2v_sgns_stage.cuh"

namespace n2v {

constexpr int MMA_ROWS = 32;            // contexts of one window, padded (window <= 16)
constexpr int MMA_LD = 132;             // row stride in floats: 128 + 4, conflict-free fragment loads
constexpr int MMA_GLD = 9;

struct MmaWarpSmem {
    float C[MMA_ROWS][MMA_LD];
    float T[8][MMA_LD];
    float G[MMA_ROWS][MMA_GLD];
};

__device__ __forceinline__ uint32_t to_tf32(float x)
{
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ void mma_tf32(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(SGNS_BLOCK, 2)
sgns_train_kernel_mma(SgnsArgs a)
{
    constexpr int FN = 5;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int32_t s_idx[SGNS_BLOCK / 32][SGNS_SMEM_TOKENS];
    __shared__ uint16_t s_pos[SGNS_BLOCK / 32][SGNS_SMEM_TOKENS];
    __shared__ uint8_t s_rw[SGNS_BLOCK / 32][SGNS_SMEM_TOKENS];
    __shared__ float s_exp[EXP_TABLE_SIZE];
    for (int i = threadIdx.x; i < EXP_TABLE_SIZE; i += blockDim.x) s_exp[i] = exp_table_entry(i);
    __syncthreads();

    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    MmaWarpSmem &S = reinterpret_cast<MmaWarpSmem *>(smem_raw)[wib];
    const WarpSentence ws{s_idx[wib], s_pos[wib], s_rw[wib]};
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = a.p.grid_warps;
    if (warp >= n_warps) return;
    const int32_t window = a.p.window;
    const uint32_t k0 = (uint32_t)a.p.seed, k1 = (uint32_t)(a.p.seed >> 32);
    const uint32_t ep8 = a.p.epoch << 8;
    const RowsFlat rows{a.syn0, a.syn1neg, 128};
    const int gq = lane >> 2, tq = lane & 3;                   // fragment coordinates: group of 4, thread in group
    unsigned long long pairs = 0, centres = 0;

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
    // rows 6, 7 of T stay zero
    for (int r = 6; r < 8; ++r) *reinterpret_cast<float4 *>(&S.T[r][lane * 4]) = make_float4(0.f, 0.f, 0.f, 0.f);

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
                int32_t j0 = i - window + ws.rw[i]; if (j0 < 0) j0 = 0;
                int32_t kend = i + window + 1 - ws.rw[i]; if (kend > n_kept) kend = n_kept;
                const int32_t n = kend - j0 - ((i >= j0 && i < kend) ? 1 : 0);
                if (n <= 0) continue;
                const int32_t centre = ws.idx[i];
                const int32_t t_cur = draw_centre(i, gs);
                int32_t tg[FN + 1];
                tg[0] = centre;
#pragma unroll
                for (int d = 0; d < FN; ++d) tg[d + 1] = __shfl_sync(0xFFFFFFFFu, t_cur, d);
                // targets used once each: a negative equal to the centre is skipped, a repeated negative is used once
                uint32_t skip = 0xC0u;
#pragma unroll
                for (int d = 1; d <= FN; ++d) {
                    if (tg[d] == centre) skip |= 1u << d;
#pragma unroll
                    for (int e = 1; e < d; ++e) if (tg[e] == tg[d]) skip |= 1u << d;
                }
                // ---- stage T (6 rows) and C (n rows; the rest of the m-tiles zero)
#pragma unroll
                for (int d = 0; d <= FN; ++d)
                    *reinterpret_cast<float4 *>(&S.T[d][lane * 4]) = ldcg4(rows.r1(tg[d]), lane);
                const int mtiles = (n + 15) >> 4;
                for (int32_t r = 0, j = j0; r < mtiles * 16; ++r) {
                    if (j == i) ++j;
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (r < n) { v = ldcg4(rows.r0(ws.idx[j]), lane); ++j; }
                    *reinterpret_cast<float4 *>(&S.C[r][lane * 4]) = v;
                }
                __syncwarp();
                // ---- F = C T^t, then G
                for (int mt = 0; mt < mtiles; ++mt) {
                    float f[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int ks = 0; ks < 16; ++ks) {
                        const float *c0 = &S.C[mt * 16 + gq][ks * 8 + tq];
                        mma_tf32(f, to_tf32(c0[0]), to_tf32(c0[8 * MMA_LD]), to_tf32(c0[4]), to_tf32(c0[8 * MMA_LD + 4]),
                                 to_tf32(S.T[gq][ks * 8 + tq]), to_tf32(S.T[gq][ks * 8 + tq + 4]));
                    }
#pragma unroll
                    for (int e = 0; e < 4; ++e) {          // f[e]: context row mt*16 + gq + 8*(e>>1), target 2*tq + (e&1)
                        const int r = mt * 16 + gq + ((e >> 1) << 3), t = 2 * tq + (e & 1);
                        float g = 0.f;
                        if (r < n && !((skip >> t) & 1u) && f[e] > -(float)MAX_EXP && f[e] < (float)MAX_EXP)
                            g = ((t == 0 ? 1.0f : 0.0f) - s_exp[(int)((f[e] + (float)MAX_EXP) * (float)(EXP_TABLE_SIZE / MAX_EXP / 2))]) * alpha;
                        S.G[r][t] = g;
                    }
                }
                __syncwarp();
                // ---- dT^t = C^t G: 8 tiles of 16 dims, k over the contexts (registers until C is no longer needed)
                float dt[8][4];
#pragma unroll
                for (int mt = 0; mt < 8; ++mt) {
                    dt[mt][0] = dt[mt][1] = dt[mt][2] = dt[mt][3] = 0.f;
                    for (int ks = 0; ks < mtiles * 2; ++ks) {
                        const float *c0 = &S.C[ks * 8 + tq][mt * 16 + gq];
                        mma_tf32(dt[mt], to_tf32(c0[0]), to_tf32(c0[8]), to_tf32(c0[4 * MMA_LD]), to_tf32(c0[4 * MMA_LD + 8]),
                                 to_tf32(S.G[ks * 8 + tq][gq]), to_tf32(S.G[ks * 8 + tq + 4][gq]));
                    }
                }
                __syncwarp();
                // ---- dC = G T, written over C
                for (int mt = 0; mt < mtiles; ++mt) {
                    const uint32_t a0 = to_tf32(S.G[mt * 16 + gq][tq]), a1 = to_tf32(S.G[mt * 16 + gq + 8][tq]),
                                   a2 = to_tf32(S.G[mt * 16 + gq][tq + 4]), a3 = to_tf32(S.G[mt * 16 + gq + 8][tq + 4]);
#pragma unroll
                    for (int nt = 0; nt < 16; ++nt) {
                        float dc[4] = {0.f, 0.f, 0.f, 0.f};
                        mma_tf32(dc, a0, a1, a2, a3, to_tf32(S.T[tq][nt * 8 + gq]), to_tf32(S.T[tq + 4][nt * 8 + gq]));
                        *reinterpret_cast<float2 *>(&S.C[mt * 16 + gq][nt * 8 + 2 * tq]) = make_float2(dc[0], dc[1]);
                        *reinterpret_cast<float2 *>(&S.C[mt * 16 + gq + 8][nt * 8 + 2 * tq]) = make_float2(dc[2], dc[3]);
                    }
                }
                __syncwarp();
                // dT over T: dt[mt][e] = dim mt*16 + gq + 8*(e>>1), target 2*tq + (e&1)
#pragma unroll
                for (int mt = 0; mt < 8; ++mt)
#pragma unroll
                    for (int e = 0; e < 4; ++e) S.T[2 * tq + (e & 1)][mt * 16 + gq + ((e >> 1) << 3)] = dt[mt][e];
                __syncwarp();
                // ---- one reduction per row
                for (int32_t r = 0, j = j0; r < n; ++r, ++j) {
                    if (j == i) ++j;
                    atomicAdd(reinterpret_cast<float4 *>(rows.r0(ws.idx[j])) + lane, *reinterpret_cast<const float4 *>(&S.C[r][lane * 4]));
                }
#pragma unroll
                for (int d = 0; d <= FN; ++d)
                    if (!((skip >> d) & 1u))
                        atomicAdd(reinterpret_cast<float4 *>(rows.r1(tg[d])) + lane, *reinterpret_cast<const float4 *>(&S.T[d][lane * 4]));
                __syncwarp();
                pairs += (unsigned long long)n;
                ++centres;
            }
            __syncwarp();
        }
    }
    if (lane == 0 && a.pairs_out && pairs) { atomicAdd(a.pairs_out, pairs); atomicAdd(a.pairs_out + 1, centres); }
}

int launch_train_mma(const SgnsArgs &a, cudaStream_t stream)
{
    if (a.p.dim != 128 || a.p.negative != 5 || a.p.window > 16 || !a.p.atomic_updates) {
        set_error("n2v_sgns_train: the tensor-core variant needs dim = 128, negative = 5, window <= 16, atomic updates");
        return N2V_EINVAL;
    }
    const size_t dyn = sizeof(MmaWarpSmem) * (SGNS_BLOCK / 32);
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(sgns_train_kernel_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn);
        if (e != cudaSuccess) { set_error("cudaFuncSetAttribute failed: %s", cudaGetErrorString(e)); return N2V_ECUDA; }
        configured = true;
    }
    const int blocks = (a.p.grid_warps + SGNS_BLOCK / 32 - 1) / (SGNS_BLOCK / 32);
    sgns_train_kernel_mma<<<blocks, SGNS_BLOCK, dyn, stream>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("n2v_sgns_train: kernel launch failed: %s", cudaGetErrorString(e)); return N2V_ECUDA; }
    return N2V_OK;
}

}  // namespace n2v
