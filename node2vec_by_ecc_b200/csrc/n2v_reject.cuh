// Rejection-sampling parameters shared by the two rejection walkers (n2v_walk.cu, n2v_walk2.cu).
#pragma once
#include <math.h>
#include <stdint.h>

namespace n2v {

struct RejectParams {
    uint32_t t_ret, t_in, t_out;   // accept iff r < t_x ; thresholds = alpha_x / B * 2^32 (saturated)
    uint32_t t_lo, t_hi;           // pre-accept below t_lo; (x != prev) pre-reject at/above t_hi
    int fold;                      // return edge folded out as an outlier (unweighted, symmetric)
    double fold_mass, bound;       // (1/p - B') and B' = max(1, 1/q)
};

static inline uint32_t to_thr(double x)   // x in [0,1] -> ceil(x * 2^32) saturated
{
    double t = ceil(x * 4294967296.0);
    if (t >= 4294967296.0) return 0xFFFFFFFFu;
    if (t <= 0.0) return 0u;
    return (uint32_t)t;
}

// alpha(prev, x) of get_alias_edge (node2vec.py:142-147): 1/p if x == prev, 1 if x is a
// neighbour of prev, 1/q otherwise; B = dartboard height.
static inline RejectParams make_reject_params(double p, double q, bool weighted, int symmetric,
                                              bool weighted_fold = false)
{
    const double a_ret = 1.0 / p, a_in = 1.0, a_out = 1.0 / q;
    RejectParams rp;
    const double b_rest = a_in > a_out ? a_in : a_out;
    // the return edge can be folded out when its weight and the row's total weight are known:
    // unit weights (K columns of area 1), or a weighted symmetric graph with strengths supplied
    rp.fold = ((!weighted || weighted_fold) && symmetric && a_ret > b_rest) ? 1 : 0;
    rp.bound = rp.fold ? b_rest : (a_ret > b_rest ? a_ret : b_rest);
    rp.fold_mass = rp.fold ? a_ret - b_rest : 0.0;
    rp.t_ret = to_thr((rp.fold ? b_rest : a_ret) / rp.bound);
    rp.t_in = to_thr(a_in / rp.bound);
    rp.t_out = to_thr(a_out / rp.bound);
    uint32_t lo = rp.t_ret < rp.t_in ? rp.t_ret : rp.t_in;
    rp.t_lo = lo < rp.t_out ? lo : rp.t_out;
    rp.t_hi = rp.t_in > rp.t_out ? rp.t_in : rp.t_out;
    return rp;
}

__device__ __forceinline__ uint32_t ceil_log2_p1(int64_t deg)   // ceil(log2(deg+1))
{
    return deg <= 0 ? 0u : (uint32_t)(64 - __clzll((unsigned long long)deg));
}

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

}  // namespace n2v
