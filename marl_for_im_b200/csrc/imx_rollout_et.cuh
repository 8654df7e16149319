// imx_rollout_et.cuh — the fused base-stock rollout with ONE THREAD PER ENVIRONMENT (runtime-specialised build, m <= 8).
//
// The lanes = nodes rollout (imx_rollout.cuh) spends one warp on 4-5 envs and resolves the stage coupling with shuffles; its
// loop is latency-bound by the per-period dependency chain with ~140 envs in flight per SM, and a divergent network runs its
// split on 2 lanes of every 6 (profiles/r2_rollout_mappings.txt).  Here a thread owns a whole env: the m nodes are an
// unrolled loop over registers, the network (parents, children, lead times, capacities) is injected as compile-time
// integer lists (-DIMX_L_<name>=v0,v1,...), the per-node cost constants sit in the kernel parameters (constant-bank
// operands, no registers), and nothing is exchanged between threads — no shuffles, no barriers, no shared memory, all 32
// lanes live, 512 envs in flight per SM.  Same arithmetic in the same order as the other kernels: bit-identical results
// (tests/test_gpu_rollout_et.py compares every env with the C oracle and with the lanes kernel on Philox demand).
// Restates dfo_func's loop (base_restock_policy.py:24-45) over the step of MAIM_env.py:330-436 / MAIM_div_env.py:441-655.
#pragma once

#include "imx_rollout.cuh"

namespace imx {

// per-node float64 constants of the env-per-thread kernels (kernel parameter -> constant bank)
struct EtNodeConsts {
    double p[8], c[8], h[8], bc[8], target[8];
};

#if defined(IMX_JIT) && defined(IMX_ET) && IMX_ET
constexpr int ET_THREADS = 128;

template <int M, int DMAX, int MAXC, bool DIV>
__global__ void __launch_bounds__(ET_THREADS) rollout_kernel_et(const __grid_constant__ StepArgs A, const __grid_constant__ RolloutArgs Rg,
                                                                const __grid_constant__ EtNodeConsts C) {
    constexpr int T = IMX_K_T, R = IMX_K_R;
    constexpr int INV_MAX[M] = {IMX_L_inv_max}, ORDER_MAX[M] = {IMX_L_order_max}, DEMAND_MAX[M] = {IMX_L_demand_max};
    constexpr int DELAY[M] = {IMX_L_delay}, INIT_INV[M] = {IMX_L_init_inv}, PIPE_OFF[M] = {IMX_L_pipe_off};
    constexpr int NCHILD[M] = {IMX_L_nchild}, RETAILER[M] = {IMX_L_retailer_idx}, BT_OFF[M] = {IMX_L_bt_off};
    constexpr int CHILDREN[M * MAXC] = {IMX_L_children};        // node j's k-th child at [j * MAXC + k] (-1: none)

    const int64_t stride = (int64_t)gridDim.x * ET_THREADS;
    for (int64_t n = (int64_t)blockIdx.x * ET_THREADS + threadIdx.x; n < A.N; n += stride) {
        int inv[M], bl[M], ou[M], cr[M], pipe[M][DMAX], bt[M][MAXC];
        double z[M], ret[M];
#pragma unroll
        for (int j = 0; j < M; ++j) {
            inv[j] = INIT_INV[j]; bl[j] = 0; ou[j] = 0; cr[j] = 0; ret[j] = 0.0;
#pragma unroll
            for (int k = 0; k < DMAX; ++k) pipe[j][k] = 0;
#pragma unroll
            for (int k = 0; k < MAXC; ++k) bt[j][k] = 0;
            z[j] = Rg.z[Rg.z_stride ? n * M + j : (int64_t)j];
        }
        const int32_t* dem = Rg.demand ? Rg.demand + n * R * T : nullptr;
        int d_odd[R];                                   // Philox: the odd period's draw of the pair
#pragma unroll
        for (int r = 0; r < R; ++r) d_odd[r] = 0;
        int err_code = 0;

        for (int t = 0; t < T; ++t) {
            int order[M], demand[M], acq[M], ship[M], incoming[M], cust[R];
            // base_stock_policy (base_restock_policy.py:12-20) + order clipping
#pragma unroll
            for (int j = 0; j < M; ++j) {
                const double om_d = (double)ORDER_MAX[j];
                const double raw = __dsub_rn(z[j], (double)(inv[j] + ou[j] - bl[j]));
                const double act = KF(std_actions) ? (raw < 0.0 ? 0.0 : (raw > om_d ? om_d : raw)) : raw;
                order[j] = decode_order(act, om_d, KF(std_actions) != 0, KF(multi) != 0, A.a, A.bma, A.inv_bma, KBMA_POW2);
            }
            // customer demand of this period: replayed trace or the Philox stream (one call serves two periods)
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (dem) {
                    cust[r] = dem[r * T + t];
                } else if ((t & 1) == 0) {
                    int d0, d1;
                    draw_demand_pair(Rg.gen, n, r, t, d0, d1);
                    cust[r] = d0;
                    d_odd[r] = d1;
                } else {
                    cust[r] = d_odd[r];
                }
            }
            // demand propagation (MAIM_env.py:351-353, MAIM_div_env.py:460-467)
#pragma unroll
            for (int j = 0; j < M; ++j) {
                if (RETAILER[j] >= 0) {
                    demand[j] = min(cust[RETAILER[j] < R ? RETAILER[j] : 0], INV_MAX[j]);
                } else if (DIV) {
                    int s = 0;
#pragma unroll
                    for (int k = 0; k < MAXC; ++k)
                        if (k < NCHILD[j]) s += order[CHILDREN[j * MAXC + k] >= 0 ? CHILDREN[j * MAXC + k] : 0];
                    demand[j] = s;
                } else {
                    demand[j] = order[j > 0 ? j - 1 : 0];
                }
            }
            // acquisition (+ noisy delay, MAIM_env.py:438-476) and shipment
#pragma unroll
            for (int j = 0; j < M; ++j) {
                int a = (t >= DELAY[j]) ? pipe[j][0] : 0;
                if (KF(has_carry)) {
                    if (Rg.noisy) {
                        a += cr[j];
                        cr[j] = 0;
                        if (t >= DELAY[j] && t < T - 1) {
                            const bool delayed = Rg.mask ? (Rg.mask[(n * T + t) * M + j] != 0)
                                                         : draw_delay(Rg.gen.seed, Rg.gen.env_offset + n, j, t, Rg.gen.episode, Rg.delay_thr);
                            if (delayed) { cr[j] = a; a = 0; }
                        }
                    }
                }
                acq[j] = a;
                ship[j] = min(bl[j] + demand[j], inv[j] + a);
            }
            // what enters each node's lead-time register: the factory's own order, the upstream shipment, or the split's share
            if constexpr (DIV) {
                incoming[0] = order[0];
#pragma unroll
                for (int j = 0; j < M; ++j) {
                    if (NCHILD[j] == 1) {
                        incoming[CHILDREN[j * MAXC] >= 0 ? CHILDREN[j * MAXC] : 0] = ship[j];
                    } else if (NCHILD[j] > 1) {
                        int od[MAXC], st[MAXC];
#pragma unroll
                        for (int k = 0; k < MAXC; ++k) od[k] = (k < NCHILD[j]) ? order[CHILDREN[j * MAXC + k] >= 0 ? CHILDREN[j * MAXC + k] : 0] : 0;
                        const int code = split_ship<MAXC>(NCHILD[j], ship[j], demand[j], bl[j], DEMAND_MAX[j], KF(wd_mult1), KF(wd_mult), od, bt[j], st);
                        if (code != 0 && err_code == 0) err_code = code;
#pragma unroll
                        for (int k = 0; k < MAXC; ++k)
                            if (k < NCHILD[j]) incoming[CHILDREN[j * MAXC + k] >= 0 ? CHILDREN[j * MAXC + k] : 0] = st[k];
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < M; ++j) incoming[j] = (j == M - 1) ? order[j] : ship[j + 1 < M ? j + 1 : j];
            }
            // state update, lead-time shift, profit (MAIM_env.py:360-384, 413-436)
            double profit[M];
#pragma unroll
            for (int j = 0; j < M; ++j) {
                int b = bl[j] + demand[j] - ship[j];
                if (KF(cap_backlog)) b = min(b, DEMAND_MAX[j]);
                ou[j] = min(max(ou[j] + order[j] - acq[j], 0), INV_MAX[j]);
                inv[j] = min(max(inv[j] + acq[j] - ship[j], 0), INV_MAX[j]);
                bl[j] = b;
#pragma unroll
                for (int k = 0; k < DMAX; ++k) {
                    const int nxt = (k + 1 < DMAX) ? pipe[j][k + 1] : 0;
                    pipe[j][k] = (k == DELAY[j] - 1) ? incoming[j] : nxt;
                }
                profit[j] = profit_of(C.p[j], C.c[j], C.h[j], C.bc[j], C.target[j], ship[j], order[j], inv[j], bl[j]);
            }
            // reward: per-agent profit, the shared mean (sequential sum from 0, / m), or np.sum for the single-agent kinds
            if (KF(multi)) {
                if (KF(independent)) {
#pragma unroll
                    for (int j = 0; j < M; ++j) {
                        ret[j] = __dadd_rn(ret[j], profit[j]);
                        if (Rg.step_reward) Rg.step_reward[((int64_t)t * A.N + n) * M + j] = profit[j];
                    }
                } else {
                    double s = 0.0;
#pragma unroll
                    for (int j = 0; j < M; ++j) s = __dadd_rn(s, profit[j]);
                    const double r = div_by_m(s, M, A.inv_m, KM_POW2);
#pragma unroll
                    for (int j = 0; j < M; ++j) {
                        ret[j] = __dadd_rn(ret[j], r);
                        if (Rg.step_reward) Rg.step_reward[((int64_t)t * A.N + n) * M + j] = r;
                    }
                }
            } else {
                double r;
                if constexpr (M < 8) {
                    r = 0.0;
#pragma unroll
                    for (int j = 0; j < M; ++j) r = __dadd_rn(r, profit[j]);
                } else {                                 // M == 8: numpy's eight-accumulator order
                    r = __dadd_rn(__dadd_rn(__dadd_rn(profit[0], profit[1]), __dadd_rn(profit[2], profit[3])),
                                  __dadd_rn(__dadd_rn(profit[4], profit[5]), __dadd_rn(profit[6], profit[7])));
                }
                ret[0] = __dadd_rn(ret[0], r);
                if (Rg.step_reward) Rg.step_reward[(int64_t)t * A.N + n] = r;
            }
        }

        if (KF(multi)) {
#pragma unroll
            for (int j = 0; j < M; ++j) Rg.ret[n * M + j] = ret[j];
        } else {
            Rg.ret[n] = ret[0];
        }
        if (Rg.write_state) {
#pragma unroll
            for (int j = 0; j < M; ++j) {
                A.inv[n * M + j] = inv[j];
                A.backlog[n * M + j] = bl[j];
                A.order_u[n * M + j] = ou[j];
                if (KF(has_carry)) A.carry[n * M + j] = cr[j];
#pragma unroll
                for (int k = 0; k < DMAX; ++k)
                    if (k < DELAY[j]) A.pipe[n * KF(L) + PIPE_OFF[j] + k] = pipe[j][k];
                if (DIV && NCHILD[j] > 1) {
#pragma unroll
                    for (int k = 0; k < MAXC; ++k)
                        if (k < NCHILD[j]) A.bt[n * KF(NB) + BT_OFF[j] + k] = bt[j][k];
                }
            }
        }
        if (DIV && err_code != 0) A.err[n] = err_code;
    }
}
#endif  // IMX_ET

}  // namespace imx
