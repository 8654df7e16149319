// imx_reset.cuh — episode start: state initialisation, initial observation, and the episode's
// demand trace / noisy-delay mask (replayed from the caller or drawn from Philox), both stored
// period-major so that the step kernel's per-period reads are contiguous across envs.
// Restates reset(): IM_env.py:164-229, MAIM_env.py:176-240, IM_div_env.py:201-302,
// MAIM_div_env.py:240-341.
#pragma once

#include "imx_step.cuh"

namespace imx {

// One thread per (env, stage) cell.  The zero part of the state is cleared by a memset before
// this kernel; here inv = init_inv and the t = 0 observation (every history / pipeline slot 0).
template <int DMAX, int PMAX>
__global__ void __launch_bounds__(256) reset_kernel(const __grid_constant__ StepArgs A, int div) {
    const int64_t cell = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= A.N * KF(m)) return;
    const int i = (int)(cell % KF(m));
    const NodeParams np = load_node(A.nodes + i);
    A.inv[cell] = np.init_inv;
    if (KHAS(obs)) {
        int pipe[DMAX], hd[PMAX], ho[PMAX];
#pragma unroll
        for (int k = 0; k < DMAX; ++k) pipe[k] = 0;
#pragma unroll
        for (int j = 0; j < PMAX; ++j) { hd[j] = 0; ho[j] = 0; }
        write_obs_row<DMAX, PMAX>(A.obs + cell * KF(O), A, np, i, KHAS(tab) ? A.tab + (size_t)i * 4 * KF(TL) : nullptr, np.init_inv, 0, 0, pipe, hd, ho, div != 0);
    }
}

// Initial observation only (used when obs is requested separately from the state reset).
// demand [N][R][T] -> [T][R][N]; one thread per (retailer row, env), env fastest.
__global__ void __launch_bounds__(256) demand_transpose_kernel(const int32_t* __restrict__ in, int32_t* __restrict__ out,
                                                               int64_t N, int R, int T) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * R) return;
    const int r = (int)(idx / N);
    const int64_t n = idx % N;
    const int32_t* src = in + (n * R + r) * T;
    for (int t = 0; t < T; ++t) out[((int64_t)t * R + r) * N + n] = src[t];
}

// Philox demand: customer_demand ~ Poisson(mu) / randint(low, high) per (env, retailer, period)
// (scipy.stats .rvs at MAIM_env.py:207,217 — same distributions, different generator).
__global__ void __launch_bounds__(256) demand_generate_kernel(int32_t* __restrict__ out, int64_t N, int R, int T, DemandGen g) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= N * R) return;
    const int r = (int)(idx / N);
    const int64_t n = idx % N;
    for (int t = 0; t < T; t += 2) {
        int d0, d1;
        draw_demand_pair(g, n, r, t, d0, d1);
        out[((int64_t)t * R + r) * N + n] = d0;
        if (t + 1 < T) out[((int64_t)(t + 1) * R + r) * N + n] = d1;
    }
}

// noisy-delay mask [N][T][m] -> [T][N][m]
__global__ void __launch_bounds__(256) mask_transpose_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out,
                                                             int64_t N, int m, int T) {
    const int64_t cell = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= N * m) return;
    const int64_t n = cell / m;
    const int i = (int)(cell % m);
    for (int t = 0; t < T; ++t) out[((int64_t)t * N + n) * m + i] = in[(n * T + t) * m + i];
}

__global__ void __launch_bounds__(256) mask_generate_kernel(uint8_t* __restrict__ out, int64_t N, int m, int T, uint64_t seed,
                                                            int64_t env_offset, uint64_t episode, double thr) {
    const int64_t cell = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= N * m) return;
    const int64_t n = cell / m;
    const int i = (int)(cell % m);
    for (int t = 0; t < T; ++t) out[((int64_t)t * N + n) * m + i] = draw_delay(seed, env_offset + n, i, t, episode, thr) ? 1 : 0;
}

}  // namespace imx
