// imx_stats.cuh — small kernels behind the rollout / evaluation entry points: the dfo_func objective and the
// deterministic episode statistics.  Host-TU only (imx_api.cu); not part of the runtime-specialised build.
#pragma once

#include "imx_device.cuh"

namespace imx {

// dfo_func's objective (base_restock_policy.py:41-45): -1 / num_periods * np.sum(prob * rewards) with prob = pmf of the
// episode's demand trace ([T], or [R, T] for a divergent network, where the product broadcasts over the retailer rows) —
// i.e. numpy's pairwise summation over the C-contiguous flattening, element j = prob[j] * rewards[j mod T].  One thread
// per env; numpy's order exactly (loops_utils.h.src, pairwise_sum): sequential below 8 elements, eight running
// accumulators up to 128, above that a recursive split at n/2 rounded down to a multiple of 8.
struct DfoTerm {
    const double* __restrict__ pmf_row;    // [R*T] of this env
    const double* __restrict__ rew;        // step rewards, element t at rew[t * stride]
    int64_t stride;
    int T;
    __device__ __forceinline__ double operator()(int j) const { return __dmul_rn(pmf_row[j], rew[(int64_t)(j % T) * stride]); }
};

// n <= 128: sequential below 8 elements, else eight running accumulators + a sequential tail
__device__ __forceinline__ double np_pairwise_block(const DfoTerm& f, int lo, int n) {
    if (n < 8) {
        double res = 0.0;
        for (int i = 0; i < n; ++i) res = __dadd_rn(res, f(lo + i));
        return res;
    }
    double r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = f(lo + j);
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = __dadd_rn(r[j], f(lo + i + j));
    }
    double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                           __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
    for (; i < n; ++i) res = __dadd_rn(res, f(lo + i));
    return res;
}

// the recursive split above 128 elements, as an explicit post-order walk (n <= 32 * 65535 -> depth <= 16)
__device__ inline double np_pairwise(const DfoTerm& f, int total) {
    struct Frame { int lo, n, stage; double left; };
    Frame st[24];
    int sp = 0;
    double ret = 0.0;
    st[sp].lo = 0; st[sp].n = total; st[sp].stage = 0; st[sp].left = 0.0; ++sp;
    while (sp > 0) {
        const int lo = st[sp - 1].lo, n = st[sp - 1].n, stage = st[sp - 1].stage;
        int n2 = n / 2;
        n2 -= n2 % 8;
        if (stage == 0) {
            if (n <= 128) { ret = np_pairwise_block(f, lo, n); --sp; }
            else { st[sp - 1].stage = 1; st[sp].lo = lo; st[sp].n = n2; st[sp].stage = 0; st[sp].left = 0.0; ++sp; }
        } else if (stage == 1) {
            st[sp - 1].left = ret; st[sp - 1].stage = 2;
            st[sp].lo = lo + n2; st[sp].n = n - n2; st[sp].stage = 0; st[sp].left = 0.0; ++sp;
        } else {
            ret = __dadd_rn(st[sp - 1].left, ret);
            --sp;
        }
    }
    return ret;
}

__global__ void __launch_bounds__(128) dfo_objective_kernel(const double* __restrict__ pmf, const double* __restrict__ step_reward,
                                                            double* __restrict__ dfo, int64_t N, int R, int T) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    DfoTerm f;
    f.pmf_row = pmf + n * (int64_t)R * T;
    f.rew = step_reward + n;
    f.stride = N;
    f.T = T;
    const double s = np_pairwise(f, R * T);
    dfo[n] = __dmul_rn(-1.0 / (double)T, s);              // "-1 / env.num_periods * np.sum(...)": (-1 / T) first, then the product
}

// Episode statistics [n, Σ total, Σ total², then per agent (Σ, Σ²)] in a fixed, deterministic
// order (the payload of the single cross-GPU all-reduce).  Two stages: STATS_BLOCKS x nstat blocks
// each reduce one statistic over a fixed slice of the envs (strided threads + shared-memory tree),
// then one block per statistic adds the STATS_BLOCKS partials in index order.  The grid is a
// constant, so the result depends only on N, never on the device.  The per-env total of a MAIM kind
// is the sum over agents in agent order.
constexpr int STATS_BLOCKS = 128;
constexpr int STATS_THREADS = 256;

__device__ __forceinline__ double stats_value(const double* __restrict__ ret, int64_t n, int cols, int q) {
    double v;
    if (q < 2) {
        v = 0.0;
        for (int c = 0; c < cols; ++c) v += ret[n * cols + c];
    } else {
        v = ret[n * cols + (q - 2) / 2];
    }
    return (q & 1) ? v * v : v;
}

__global__ void __launch_bounds__(STATS_THREADS) return_stats_partial_kernel(const double* __restrict__ ret, double* __restrict__ partial,
                                                                             int64_t N, int cols) {
    __shared__ double red[STATS_THREADS];
    const int q = blockIdx.y;
    const int64_t per_block = (N + STATS_BLOCKS - 1) / STATS_BLOCKS;
    const int64_t lo = (int64_t)blockIdx.x * per_block;
    const int64_t hi = lo + per_block < N ? lo + per_block : N;
    double acc = 0.0;
    for (int64_t n = lo + threadIdx.x; n < hi; n += STATS_THREADS) acc += stats_value(ret, n, cols, q);
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = STATS_THREADS / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[q * STATS_BLOCKS + blockIdx.x] = red[0];
}

// One warp per statistic: fixed-shape tree over the STATS_BLOCKS partials (deterministic).
// accumulate != 0: stats += result (statistics of an evaluation batch build up on the device).
__global__ void __launch_bounds__(32) return_stats_final_kernel(const double* __restrict__ partial, double* __restrict__ stats, int64_t N,
                                                                int nstat, int accumulate) {
    const int q = blockIdx.x;
    const int lane = threadIdx.x;
    double acc = 0.0;
    for (int b = lane; b < STATS_BLOCKS; b += 32) acc += partial[q * STATS_BLOCKS + b];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, off);
    if (lane == 0) {
        stats[1 + q] = accumulate ? stats[1 + q] + acc : acc;
        if (q == 0) stats[0] = accumulate ? stats[0] + (double)N : (double)N;
    }
}

// Episode return per (env, agent) = sum over periods of the step rewards, added in period order like
// the host loops of the reference ("reward += r", inv_management.py:223-231).  One thread per cell;
// each period is one coalesced row of step_reward [T][cells].
__global__ void __launch_bounds__(256) episode_return_kernel(const double* __restrict__ step_reward, double* __restrict__ ret, int64_t cells,
                                                             int T) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cells) return;
    double acc = 0.0;
    for (int t = 0; t < T; ++t) acc += step_reward[(int64_t)t * cells + c];
    ret[c] = acc;
}

}  // namespace imx
