// imx_rollout.cuh — fused K-period base-stock rollout: a whole episode per environment in ONE
// kernel, state in registers (the lead-time pipeline is a register shift register), demand from
// the counter-based Philox stream or a replayed trace, the order-up-to policy evaluated in-lane.
// Restates the loop of dfo_func (base_restock_policy.py:30-45) with base_stock_policy (:4-21) and
// the env step (same arithmetic as imx_step.cuh) inlined; per-env traffic is the demand trace in
// and the returns out (≈1 byte per agent-step), so this kernel is issue-bound, not HBM-bound.
#pragma once

#include "imx_step.cuh"

namespace imx {

constexpr int ROLLOUT_THREADS = 128;

struct RolloutArgs {
    const double* __restrict__ z;          // base-stock levels [m] or [N][m]
    int32_t z_stride;                      // 0 or m
    int32_t write_state;
    const int32_t* __restrict__ demand;    // replayed [N][R][T] or nullptr → Philox
    int32_t coop_demand;                   // Philox draws are produced cooperatively by the tile's lanes into shared memory
    int32_t pad_r;
    double* __restrict__ ret;              // [N] (IM kinds) / [N][m] (MAIM kinds)
    double* __restrict__ step_reward;      // [T][N] / [T][N][m] or nullptr (the dfo objective is a second kernel over this array)
    const uint8_t* __restrict__ mask;      // replayed noisy-delay outcomes [N][T][m] or nullptr -> Philox (same draws as imx_reset)
    int32_t noisy;                         // this episode runs with noisy delays (handles created with noisy_delay = 1 only)
    int32_t pad_n;
    double delay_thr;                      // Philox mask: P(delay) per eligible (stage, period)
    DemandGen gen;
};

#if defined(IMX_JIT) && defined(IMX_ROLLOUT_MAXNREG)
#define IMX_ROLLOUT_BOUNDS __maxnreg__(IMX_ROLLOUT_MAXNREG)      /* experiment knob of the specialised build (host: IMX_ROLLOUT_MAXNREG) */
#else
#define IMX_ROLLOUT_BOUNDS __launch_bounds__(ROLLOUT_THREADS)
#endif
template <int M_PAD, int DMAX, int MAXC, bool DIV>
__global__ void IMX_ROLLOUT_BOUNDS rollout_kernel(const __grid_constant__ StepArgs A,
                                                                  const __grid_constant__ RolloutArgs Rg) {
    constexpr int EPW = 32 / M_PAD;
    extern __shared__ int32_t s_draws[];            // [warps][EPW][R][T_even] when Rg.coop_demand
    __shared__ __align__(16) double s_profit[ROLLOUT_THREADS / 32][2][32];   // per warp, per period parity: the lanes' profits
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int i = lane % M_PAD;
    const int sub = lane / M_PAD;
    const int tbase = lane - i;
    const int m = KF(m), T = KF(T);
    const bool stage_ok = i < m && sub < EPW;         // M_PAD is the tile width (exactly m in the specialised build)
    const int T_even = (T + 1) & ~1;
    int32_t* my_draws = s_draws + ((size_t)(warp * EPW + sub) * KF(R)) * T_even;   // this env's [R][T_even] block

    const NodeParams np = load_node(A.nodes + (stage_ok ? i : 0));
    int child_lane[MAXC];
    if constexpr (DIV) {
#pragma unroll
        for (int k = 0; k < MAXC; ++k)
            child_lane[k] = stage_ok ? child_lane_of(np, k) : -1;
    }
    const bool is_last = (i == m - 1);
    const int delay_m1 = np.delay - 1;
    const double om_d = (double)np.order_max;

    const int64_t warps_in_grid = (int64_t)gridDim.x * (ROLLOUT_THREADS / 32);
    const int64_t n_warp_tiles = (A.N + EPW - 1) / EPW;
    for (int64_t wt = (int64_t)blockIdx.x * (ROLLOUT_THREADS / 32) + warp; wt < n_warp_tiles; wt += warps_in_grid) {
        const int64_t n = wt * EPW + sub;
        const bool ok = stage_ok && n < A.N;
        const int64_t cell = n * m + i;

        // reset state (MAIM_env.py:232-235): inv = init_inv, everything else 0
        int inv = np.init_inv, backlog = 0, order_u = 0, carry = 0;
        int pipe[DMAX];
        int bt[MAXC];
#pragma unroll
        for (int k = 0; k < DMAX; ++k) pipe[k] = 0;
#pragma unroll
        for (int k = 0; k < MAXC; ++k) bt[k] = 0;
        const double z = ok ? Rg.z[Rg.z_stride ? cell : (int64_t)i] : 0.0;
        const int32_t* dem_row = (ok && Rg.demand && np.retailer_idx >= 0) ? Rg.demand + (n * KF(R) + np.retailer_idx) * T : nullptr;
        if (Rg.coop_demand) {
            // every lane of the tile draws a share of the episode's (retailer, period-pair) demands
            __syncwarp();                           // previous env's reads are done
            if (n < A.N && sub < EPW) {
                const int pairs = KF(R) * (T_even >> 1);
                for (int q = i; q < pairs; q += M_PAD) {
                    const int r = q / (T_even >> 1), t2 = (q % (T_even >> 1)) * 2;
                    int d0, d1;
                    draw_demand_pair(Rg.gen, n, r, t2, d0, d1);
                    my_draws[r * T_even + t2] = d0;
                    my_draws[r * T_even + t2 + 1] = d1;
                }
            }
            __syncwarp();
        }
        double ret = 0.0;                 // "dfo_reward = 0; dfo_reward += r" (inv_management.py:223-231)
        int err_code = 0;

        for (int t = 0; t < T; ++t) {
            // base_stock_policy: z - (inv + order_u - backlog), clipped to [0, order_max]  (base_restock_policy.py:12-20)
            // (the three integers are exact in float64, so is their sum: one conversion instead of three + two additions)
            const double raw = __dsub_rn(z, (double)(inv + order_u - backlog));
            // with raw (non-standardised) actions the policy's clip to [0, order_max] is subsumed by the env's own clip
            const double act = KF(std_actions) ? (raw < 0.0 ? 0.0 : (raw > om_d ? om_d : raw)) : raw;   // z is finite
            const int order = ok ? decode_order(act, om_d, KF(std_actions) != 0, KF(multi) != 0, A.a, A.bma, A.inv_bma, KBMA_POW2) : 0;

            int cust = 0;
            if (ok && np.retailer_idx >= 0)
                cust = dem_row ? dem_row[t] : (Rg.coop_demand ? my_draws[np.retailer_idx * T_even + t] : draw_demand(Rg.gen, n, np.retailer_idx, t));

            int demand;
            int od[MAXC];
            if constexpr (DIV) {
                int s = 0;
#pragma unroll
                for (int k = 0; k < MAXC; ++k) {
                    od[k] = 0;
                    if (k < KF(maxc)) {
                        const int v = __shfl_sync(0xffffffffu, order, tbase + (child_lane[k] < 0 ? 0 : child_lane[k]));
                        od[k] = child_lane[k] < 0 ? 0 : v;
                        s += od[k];
                    }
                }
                demand = (np.retailer_idx >= 0) ? min(cust, np.inv_max) : s;
            } else {
                const int down = __shfl_up_sync(0xffffffffu, order, 1);
                demand = (i == 0) ? min(cust, np.inv_max) : down;
            }

            // update_acquisition (MAIM_env.py:438-476): head of the lead-time register, plus what a noisy delay held back
            int acq = (t >= np.delay) ? pipe[0] : 0;
            if (KF(has_carry)) {
                if (Rg.noisy) {                              // sticky noisy_delay (MAIM_env.py:192-194): dfo_func after a noisy reset
                    acq += carry;
                    carry = 0;
                    if (ok && t >= np.delay && t < T - 1) {
                        const bool delayed = Rg.mask ? (Rg.mask[(n * T + t) * m + i] != 0)
                                                     : draw_delay(Rg.gen.seed, Rg.gen.env_offset + n, i, t, Rg.gen.episode, Rg.delay_thr);
                        if (delayed) { carry = acq; acq = 0; }
                    }
                }
            }
            const int ship = min(backlog + demand, inv + acq);
            int incoming;
            if constexpr (DIV) {
                int st[MAXC];
#pragma unroll
                for (int k = 0; k < MAXC; ++k) st[k] = 0;
                if (ok && np.nchild == 1) st[0] = ship;
                if (ok && np.nchild > 1) {
                    const int code = split_ship<MAXC>(np.nchild, ship, demand, backlog, np.demand_max, KF(wd_mult1), KF(wd_mult), od, bt, st);
                    if (code != 0 && err_code == 0) err_code = code;
                }
                incoming = order;
#pragma unroll
                for (int k = 0; k < MAXC; ++k) {
                    if (k < KF(maxc)) {
                        const int v = __shfl_sync(0xffffffffu, st[k], tbase + (np.parent < 0 ? 0 : np.parent));
                        if (np.parent >= 0 && np.child_slot == k) incoming = v;
                    }
                }
            } else {
                const int up = __shfl_down_sync(0xffffffffu, ship, 1);
                incoming = is_last ? order : up;
            }

            int backlog_new = backlog + demand - ship;
            if (KF(cap_backlog)) backlog_new = min(backlog_new, np.demand_max);
            order_u = min(max(order_u + order - acq, 0), np.inv_max);
            inv = min(max(inv + acq - ship, 0), np.inv_max);
            backlog = backlog_new;
#pragma unroll
            for (int k = 0; k < DMAX; ++k) {
                const int nxt = (k + 1 < DMAX) ? pipe[k + 1] : 0;
                pipe[k] = (k == delay_m1) ? incoming : nxt;
            }

            const double profit = ok ? profit_of(np.p, np.c, np.h, np.bc, np.target, ship, order, inv, backlog) : 0.0;
            double r;
            if (KF(multi)) {
                if (KF(independent)) {
                    r = profit;
                } else if constexpr ((M_PAD & (M_PAD - 1)) != 0) {   // m-wide tiles (m = 6: measured faster with shuffles)
                    r = div_by_m(tile_seq_sum<M_PAD>(profit, m, tbase), m, A.inv_m, KM_POW2);
                } else {                                     // shared reward: the tile's profits through shared memory
                    double* wbuf = s_profit[warp][t & 1];
                    wbuf[lane] = profit;
                    __syncwarp();                            // (two buffers: a lane can run at most one period ahead)
                    r = div_by_m(tile_seq_sum_smem<M_PAD>(wbuf, m, tbase), m, A.inv_m, KM_POW2);
                }
            }
            else r = tile_np_sum<M_PAD>(profit, m, tbase);
            ret = __dadd_rn(ret, r);
            if (ok && Rg.step_reward) {
                if (KF(multi)) Rg.step_reward[(int64_t)t * A.N * m + cell] = r;
                else if (i == 0) Rg.step_reward[(int64_t)t * A.N + n] = r;
            }
        }
        __syncwarp();                                    // the next episode's period 0 reuses a profit buffer (odd T: the last one read)

        if (ok) {
            if (KF(multi)) Rg.ret[cell] = ret;
            else if (i == 0) Rg.ret[n] = ret;
            if (Rg.write_state) {
                A.inv[cell] = inv;
                A.backlog[cell] = backlog;
                A.order_u[cell] = order_u;
                int32_t* pp = A.pipe + n * KF(L) + np.pipe_off;
#pragma unroll
                for (int k = 0; k < DMAX; ++k)
                    if (k < np.delay) pp[k] = pipe[k];
                if (KF(has_carry)) A.carry[cell] = carry;
                if constexpr (DIV) {
                    if (np.bt_off >= 0) {
#pragma unroll
                        for (int k = 0; k < MAXC; ++k)
                            if (k < np.nchild) A.bt[n * KF(NB) + np.bt_off + k] = bt[k];
                    }
                }
            }
            if constexpr (DIV) {
                if (err_code != 0) A.err[n] = err_code;
            }
        }
    }
}


}  // namespace imx
