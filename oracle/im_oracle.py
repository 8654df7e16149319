"""TEST INFRASTRUCTURE — CPU restatement (oracle) of the MARL-for-IM environment step.

This file is the checker, not the product: only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s CPU-baseline legs may import it.  The product path
(``marl_for_im_b200``) never does and has no CPU fallback.

It restates, as a *state machine* (one small record per env instead of the
reference's ``[T+1, m]`` history arrays), the algorithm of

    environments/IM_env.py        (kind "IM")        single-agent serial chain
    environments/MAIM_env.py      (kind "MAIM")      multi-agent serial chain
    environments/IM_div_env.py    (kind "IM_div")    single-agent divergent tree
    environments/MAIM_div_env.py  (kind "MAIM_div")  multi-agent divergent tree
    base_restock_policy.py        (base_stock_policy / dfo_func)

Every function cites the reference lines it follows (paths relative to the
reference root).  Parity pinning: the reference has no tests or golden vectors of
its own, so this oracle is pinned by (a) running the unmodified
reference in the build container through ``oracle/ref_import.py`` — see
``tests/test_oracle_vs_reference.py`` — and (b) the committed fixtures under
``tests/golden/`` that were generated from the reference by
``tests/golden/make_golden.py``.

Pure-Python scalar loops: meant for small cases (thousands of env-steps).  The C
restatement ``oracle/imx_oracle.c`` covers full-size batches.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import numpy as np

KINDS = ("IM", "MAIM", "IM_div", "MAIM_div")


class OracleWatchdog(Exception):
    """Mirrors the reference's ``raise Exception("Infinite Loop k")`` in the divergent split."""

    def __init__(self, code: int, node: int):
        super().__init__(f"Infinite Loop {code} at node {node}")
        self.code, self.node = code, node


# --------------------------------------------------------------------------------------
# small helpers
# --------------------------------------------------------------------------------------
def np_sum_order(vals: Sequence[float]) -> float:
    """Summation order of ``np.sum`` on a contiguous float64 vector (numpy's pairwise_sum,
    numpy/core/src/umath/loops_utils.h.src: sequential for n < 8, eight running accumulators up to
    128 elements, above that a recursive split at n/2 rounded down to a multiple of 8).  Used for
    IM_env.py:372, IM_div_env.py:561 and the dfo objective (base_restock_policy.py:45).  Checked
    against numpy 2.3.5 for n = 2..33 and for n up to 70 000 in tests."""
    n = len(vals)
    if n < 8:
        res = 0.0
        for v in vals:
            res += v
        return res
    if n > 128:
        n2 = n // 2
        n2 -= n2 % 8
        return np_sum_order(vals[:n2]) + np_sum_order(vals[n2:])
    r = [vals[j] for j in range(8)]
    i = 8
    while i < n - (n % 8):
        for j in range(8):
            r[j] += vals[i + j]
        i += 8
    res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
    while i < n:
        res += vals[i]
        i += 1
    return res


def rescale(v: float, vmax: float, a: float, b: float) -> float:
    """rescale(val, 0, max, A, B) — MAIM_env.py:497-507, IM_env.py:435-444 (min_val is always 0)."""
    return a + (((v - 0.0) * (b - a)) / (vmax - 0.0))


def rev_scale(x: float, vmax: float, a: float, b: float) -> float:
    """rev_scale(val, 0, max, A, B) — MAIM_env.py:509-519, IM_env.py:446-456."""
    return (((x - a) * (vmax - 0.0)) / (b - a)) + 0.0


def _rint(x: float) -> float:
    """np.round(x, 0): round-half-to-even (MAIM_env.py:344)."""
    return float(np.rint(x))


def _as_int_vec(x, n: int, name: str) -> List[int]:
    arr = np.asarray(x, dtype=np.float64).reshape(-1)
    if arr.size != n:
        raise ValueError(f"{name}: expected {n} entries, got {arr.size}")
    if not np.all(arr == np.rint(arr)):
        raise ValueError(f"{name} must be integral")
    return [int(v) for v in arr]


def _as_f_vec(x, n: int, name: str) -> List[float]:
    arr = np.asarray(x, dtype=np.float64).reshape(-1)
    if arr.size != n:
        raise ValueError(f"{name}: expected {n} entries, got {arr.size}")
    return [float(v) for v in arr]


# --------------------------------------------------------------------------------------
# topology — utils.py:94-130
# --------------------------------------------------------------------------------------
def topology(connections: Dict[int, List[int]], num_nodes: int):
    """parent / children / depth / retailers of a tree given as {parent: [children]}.

    utils.py:94-102 (adjacency), :105-120 (depth = edges up to node 0), :124-130 (leaves);
    parent(i) = first row with network[row][i] == 1 (MAIM_div_env.py:33-35)."""
    children = [list(connections.get(i, []) or []) for i in range(num_nodes)]
    parent = [-1] * num_nodes
    for p in range(num_nodes):
        for c in children[p]:
            if c < p:
                raise ValueError("Downstream node cannot have a smaller index number than upstream node")  # utils.py:87-92
            if parent[c] == -1 or p < parent[c]:
                parent[c] = p
    depth = [0] * num_nodes
    for i in range(1, num_nodes):
        d, node = 0, i
        while node != 0:
            if parent[node] < 0:
                raise ValueError(f"node {i} is not connected to node 0")
            node = parent[node]
            d += 1
        depth[i] = d
    retailers = [i for i in range(num_nodes) if not children[i]]
    return parent, children, depth, retailers


# --------------------------------------------------------------------------------------
# the oracle env
# --------------------------------------------------------------------------------------
class OracleEnv:
    """One environment instance; ``kind`` selects which reference class is restated."""

    def __init__(self, kind: str, config: dict):
        if kind not in KINDS:
            raise ValueError(kind)
        cfg = dict(config)
        self.kind = kind
        self.div = kind.endswith("_div")
        self.multi = kind.startswith("MAIM")
        g = cfg.get

        self.T = int(g("num_periods", 50))                                   # MAIM_env.py:13
        if self.div:
            m = int(g("num_nodes", 3))                                        # MAIM_div_env.py:19
            self.connections = {int(k): list(v) for k, v in g("connections", {0: [1], 1: [2], 2: []}).items()}
            self.parent, self.children, self.depth, self.retailers = topology(self.connections, m)
        else:
            m = int(g("num_stages", 3))                                       # MAIM_env.py:17
            self.retailers = [0]
        self.m = m
        self.R = len(self.retailers)

        # per-class defaults: IM_env.py:17,39,50  MAIM_env.py:29,30,57  IM_div_env.py:31,63,74  MAIM_div_env.py:42,43,82
        init_default = 100 if self.multi else 20
        inv_max_default = 200 if kind == "MAIM" else 100
        target_default = 0 if kind == "MAIM_div" else 10
        self.init_inv = _as_int_vec(g("init_inv", np.ones(m) * init_default), m, "init_inv")
        self.inv_target = _as_f_vec(g("inv_target", np.ones(m) * target_default), m, "inv_target")
        self.inv_max = _as_int_vec(g("inv_max", np.ones(m) * inv_max_default), m, "inv_max")
        self.delay = _as_int_vec(g("delay", np.ones(m)), m, "delay")
        if min(self.delay) < 1:
            raise ValueError("delay must be >= 1 for every stage (delay = 0 is outside the supported contract)")
        self.D = max(self.delay)                                             # MAIM_env.py:36
        self.stock_cost = _as_f_vec(g("stock_cost", np.ones(m) * 0.5), m, "stock_cost")
        self.backlog_cost = _as_f_vec(g("backlog_cost", np.ones(m)), m, "backlog_cost")

        # flags
        if kind == "MAIM_div":                                               # quirk 9: flags ignored, always standardised
            self.std_state = self.std_actions = True
        else:
            self.std_state = bool(g("standardise_state", True))
            self.std_actions = bool(g("standardise_actions", True))
        if kind == "IM_div":                                                 # IM_div_env.py:35-36
            self.a, self.b = -1.0, 1.0
        else:
            self.a, self.b = float(g("a", -1)), float(g("b", 1))
        self.td = bool(g("time_dependency", False))
        self.pa = bool(g("prev_actions", False))
        self.pd = bool(g("prev_demand", False))
        self.P = int(g("prev_length", 1))
        self.independent = bool(g("independent", True))                      # MAIM_env.py:16
        self.share_network = bool(g("share_network", False)) if kind == "MAIM_div" else False
        if self.multi and (not self.td) and self.pa and (not self.pd):
            raise Exception("Not Implemented")                               # MAIM_env.py:135-136, MAIM_div_env.py:164-165

        # prices / capacities
        if self.div:
            self.p = [float(self.depth[i] + 2) for i in range(m)]            # MAIM_div_env.py:55-61
            self.c = [float(self.depth[i] + 1) for i in range(m)]
            order_max = [self.inv_max[0]] + [self.inv_max[self.parent[i]] for i in range(1, m)]   # :83-86
        else:
            price = _as_f_vec(g("price", np.flip(np.arange(m + 1) + 1)), m + 1, "price")            # MAIM_env.py:41
            for i in range(m):
                assert price[i] > price[i + 1]                               # MAIM_env.py:167-168
            self.p = price[:m]
            self.c = price[1:]
            order_max = [self.inv_max[i + 1] for i in range(m - 1)] + [self.inv_max[m - 1]]        # :58-61
        self.order_max = _as_int_vec(g("order_max", order_max), m, "order_max")
        if self.div:
            assert self.order_max[0] <= self.inv_max[0]                      # MAIM_div_env.py:235
            self.demand_max = list(self.inv_max)                             # :91-99
            for i in range(m):
                s = sum(self.order_max[c] for c in self.children[i])
                if s > self.demand_max[i]:
                    self.demand_max[i] = s
            self.split_nodes = [i for i in range(m) if len(self.children[i]) > 1]
        else:
            assert self.order_max[m - 1] <= self.inv_max[m - 1]              # MAIM_env.py:171
            self.demand_max = list(self.inv_max)
            self.split_nodes = []

        # observation length
        self.O = 3 + (self.P if self.pd else 0) + (self.P if self.pa else 0) + (self.D if self.td else 0) \
            + (1 if self.share_network else 0)

        self.noisy_delay = False
        self.reset(np.zeros((self.R, self.T), dtype=np.int64) if self.div else np.zeros(self.T, dtype=np.int64))

    # ------------------------------------------------------------------
    def reset(self, customer_demand, delay_mask: Optional[np.ndarray] = None):
        """reset(customer_demand=...) — MAIM_env.py:176-240 / MAIM_div_env.py:240-341.

        ``customer_demand``: [T] (serial) or [R, T] (divergent).  ``delay_mask[t, i]`` (bool)
        replays the noisy-delay Bernoulli draw ``u <= threshold`` made for stage i in
        period t (MAIM_env.py:449-452); None = no noisy delay."""
        m, D, P = self.m, self.D, self.P
        d = np.asarray(customer_demand)
        self.demand_trace = d.reshape(self.R, -1) if self.div else d.reshape(1, -1)
        self.delay_mask = None if delay_mask is None else np.asarray(delay_mask, dtype=bool)
        self.noisy_delay = delay_mask is not None
        self.t = 0
        self.inv = list(self.init_inv)
        self.backlog = [0] * m
        self.order_u = [0] * m
        self.pipe = [[0] * D for _ in range(m)]
        self.hist_d = [[0] * P for _ in range(m)]
        self.hist_o = [[0] * P for _ in range(m)]
        self.carry = [0] * m
        self.backlog_to = {i: [0] * len(self.children[i]) for i in self.split_nodes}
        return self.observe()

    # ------------------------------------------------------------------
    def observe(self) -> np.ndarray:
        """_update_state — IM_env.py:231-285, MAIM_env.py:242-328, IM_div_env.py:304-359,
        MAIM_div_env.py:343-439.  Returns [m, O] float64 (row i = agent i's vector)."""
        m, D, P, a, b, t = self.m, self.D, self.P, self.a, self.b, self.t
        obs = np.zeros((m, self.O))
        for i in range(m):
            inv_max, order_max = float(self.inv_max[i]), float(self.order_max[i])
            dem_max = float(self.demand_max[i]) if self.div else inv_max
            # which maximum scales order_u: MAIM_env.py:300 / MAIM_div_env.py:415 use order_max,
            # IM_env.py:265 / IM_div_env.py:339 use inv_max
            ou_max = order_max if self.multi else inv_max
            row = obs[i]
            if self.std_state:
                row[0] = rescale(self.inv[i], inv_max, a, b)
                row[1] = rescale(self.backlog[i], dem_max, a, b)
                row[2] = rescale(self.order_u[i], ou_max, a, b)
            else:
                row[0], row[1], row[2] = self.inv[i], self.backlog[i], self.order_u[i]
            # history values: slot j = value at period t-1-j, 0 if j >= t (MAIM_env.py:272-274)
            dh = [rescale(self.hist_d[i][j] if j < t else 0, dem_max, a, b) for j in range(P)]
            oh = [rescale(self.hist_o[i][j] if j < t else 0, order_max, a, b) for j in range(P)]
            raw_pipe = [self.pipe[i][k] if t >= 1 else 0 for k in range(D)]
            if self.kind == "MAIM_div":                                      # MAIM_div_env.py:408-411
                pv = [rescale(min(v, 2 * self.inv_max[i]), 2.0 * inv_max, a, b) for v in raw_pipe]
            else:
                pv = [rescale(v, inv_max, a, b) for v in raw_pipe]

            if self.multi and not self.std_state:
                # MAIM_env.py:319-324 (quirk 13): raw pipe at [3:3+D] regardless of history offsets
                if t >= 1 and self.td:
                    row[3:3 + D] = raw_pipe
                continue
            if not self.multi and not self.std_state:
                pv = raw_pipe                                                # IM_env.py:257-260 only rescales if standardised
            k = 3
            if self.multi and self.pd and not self.pa and not self.td:
                k += P                                                       # quirk 2: slot exists, never written
            else:
                if self.pd:
                    row[k:k + P] = dh
                    k += P
                if self.pa:
                    row[k:k + P] = oh
                    k += P
            if self.td:
                row[k:k + D] = pv
                k += D
            if self.share_network:                                           # MAIM_div_env.py:434-435
                row[self.O - 1] = rescale(i, float(self.m), a, b)
        return obs

    # ------------------------------------------------------------------
    def decode_orders(self, actions) -> List[int]:
        """Order clipping — IM_env.py:298-302, MAIM_env.py:340-347, IM_div_env.py:372-376,
        MAIM_div_env.py:451-456."""
        out = []
        for i in range(self.m):
            x = float(np.asarray(actions[i], dtype=np.float64).reshape(-1)[0])
            om = float(self.order_max[i])
            if self.std_actions:
                x = rev_scale(x, om, self.a, self.b)
            if self.multi:
                # round, .astype(int), then clip (MAIM_env.py:344-347).  The reference runs on x86-64, where the float64 ->
                # int64 conversion of NaN, +-inf and anything outside [-2^63, 2^63) yields INT64_MIN, which then clips to 0
                x = _rint(x)
                if not (-9.223372036854775808e18 <= x < 9.223372036854775808e18):
                    x = -9.223372036854775808e18
                x = min(max(x, 0.0), om)
            else:
                # clip, then round, then .astype(int) (IM_env.py:300-302): +-inf clip like any value; a NaN survives the clip
                # and becomes an order of INT64_MIN that corrupts the reference's float state — outside the parity domain
                if x != x:
                    raise ValueError("NaN action on a single-agent env: the reference's state becomes INT64_MIN garbage (IM_env.py:300)")
                x = _rint(min(max(x, 0.0), om))
            out.append(int(x))
        return out

    # ------------------------------------------------------------------
    def _split(self, i: int, ship: int, demand: int, backlog: int, orders: List[int]) -> List[int]:
        """Divergent split for node i — MAIM_div_env.py:476-579 / IM_div_env.py:396-502.
        Returns ship_to per child (in listed order); updates the signed ledger backlog_to[i]."""
        C = self.children[i]
        nc = len(C)
        if nc == 1:
            return [ship]
        dm = self.demand_max[i]
        th = (2 * dm, dm, dm, dm) if self.multi else (4 * dm, 2 * dm, 2 * dm, 2 * dm)
        bt = self.backlog_to[i]
        st = [0] * nc
        amt = ship
        od = [orders[c] for c in C]

        def loop1(code, limit):
            nonlocal amt
            cnt = 0
            while sum(bt) > 0 and amt > 0:
                for k in range(nc):
                    if bt[k] > 0:
                        st[k] += 1
                        bt[k] -= 1
                        amt -= 1
                cnt += 1
                if cnt > limit:
                    raise OracleWatchdog(code, i)

        if ship >= demand:
            if backlog > 0:
                loop1(1, th[0])
                if amt > 0 and demand > 0:
                    out = list(od)
                    cnt = 0
                    while amt > 0 and sum(out) > 0:
                        for k in range(nc):
                            if out[k] > 0:
                                st[k] += 1
                                out[k] -= 1
                                amt -= 1
                        cnt += 1
                        if cnt > th[1]:
                            raise OracleWatchdog(2, i)
                    for k in range(nc):
                        bt[k] += out[k]
            else:
                for k in range(nc):
                    st[k] += od[k]
        else:
            if backlog > 0:
                loop1(3, th[2])
            else:
                cnt = 0
                while amt > 0:
                    for k in range(nc):
                        if st[k] < od[k] + bt[k]:
                            st[k] += 1
                            amt -= 1
                    cnt += 1
                    if cnt > th[3]:
                        raise OracleWatchdog(4, i)
            for k in range(nc):
                bt[k] += od[k] - st[k]
        return st

    # ------------------------------------------------------------------
    def step(self, actions):
        """step — IM_env.py:287-360, MAIM_env.py:330-411, IM_div_env.py:361-549, MAIM_div_env.py:441-630.

        Returns (obs [m, O], reward, done, info) with reward a float (IM kinds) or a
        length-m float64 array (MAIM kinds) and info a dict of length-m arrays."""
        m, t, D, P = self.m, self.t, self.D, self.P
        if t >= self.T:
            raise IndexError("step() past the end of the episode")
        order = self.decode_orders(actions)

        # demand propagation — MAIM_env.py:351-353 ; MAIM_div_env.py:460-467
        demand = [0] * m
        if self.div:
            for k, r in enumerate(self.retailers):
                demand[r] = min(int(self.demand_trace[k][t]), self.inv_max[r])
            for i in range(m):
                if self.children[i]:
                    demand[i] = sum(order[c] for c in self.children[i])
        else:
            demand[0] = min(int(self.demand_trace[0][t]), self.inv_max[0])
            for i in range(1, m):
                demand[i] = order[i - 1]

        # acquisition (pipeline head) + optional replayed noisy delay —
        # MAIM_env.py:438-476 ; MAIM_div_env.py:657-695.  (The reference's draw order —
        # factory first — only matters for how the caller builds delay_mask.)
        acq = [0] * m
        new_carry = [0] * m
        for i in range(m):
            a_i = self.carry[i]
            if t >= self.delay[i]:
                a_i += self.pipe[i][0]
                if self.noisy_delay and self.delay_mask[t][i] and t < self.T - 1:
                    new_carry[i] = a_i
                    a_i = 0
            acq[i] = a_i

        # shipment — MAIM_env.py:360
        ship = [min(self.backlog[i] + demand[i], self.inv[i] + acq[i]) for i in range(m)]

        # what enters each pipeline this period — MAIM_env.py:491-495 ; MAIM_div_env.py:710-715
        incoming = [0] * m
        if self.div:
            ship_to = {}
            for i in range(m):
                if self.children[i]:
                    st = self._split(i, ship[i], demand[i], self.backlog[i], order)
                    for k, c in enumerate(self.children[i]):
                        if self.parent[c] == i:
                            ship_to[c] = st[k]
            incoming[0] = order[0]
            for i in range(1, m):
                incoming[i] = ship_to[i]
        else:
            incoming[m - 1] = order[m - 1]
            for i in range(m - 1):
                incoming[i] = ship[i + 1]

        # backlog / order_u / inv — MAIM_env.py:363-384 ; MAIM_div_env.py:582-603
        cap_backlog = True if self.kind == "MAIM_div" else self.std_state
        for i in range(m):
            bl = self.backlog[i] + demand[i] - ship[i]
            if cap_backlog:
                bl = min(bl, self.demand_max[i] if self.div else self.inv_max[i])
            ou = min(max(self.order_u[i] + order[i] - acq[i], 0), self.inv_max[i])
            iv = min(max(self.inv[i] + acq[i] - ship[i], 0), self.inv_max[i])
            self.backlog[i], self.order_u[i], self.inv[i] = bl, ou, iv
            # pipeline shift + insert — MAIM_env.py:486-495
            self.pipe[i] = self.pipe[i][1:] + [0]
            self.pipe[i][self.delay[i] - 1] = incoming[i]
            self.hist_d[i] = ([demand[i]] + self.hist_d[i])[:P]
            self.hist_o[i] = ([order[i]] + self.hist_o[i])[:P]
        self.carry = new_carry

        # rewards — MAIM_env.py:413-436, IM_env.py:362-374, MAIM_div_env.py:632-655, IM_div_env.py:551-563
        profit = []
        for i in range(m):
            pr = self.p[i] * float(ship[i]) - self.c[i] * float(order[i]) \
                - self.stock_cost[i] * abs(float(self.inv[i]) - self.inv_target[i]) \
                - self.backlog_cost[i] * float(self.backlog[i])
            profit.append(pr)
        if self.multi:
            if self.independent:
                reward = np.array(profit, dtype=np.float64)
            else:
                s = 0
                for pr in profit:
                    s += pr
                reward = np.full(m, s / m, dtype=np.float64)
        else:
            reward = float(np_sum_order(profit))

        self.t += 1
        info = {
            # IM kinds report the pre-increment period (IM_env.py:346), MAIM kinds post-increment (MAIM_env.py:402)
            "period": self.t if self.multi else self.t - 1,
            "demand": np.array(demand, dtype=np.float64),
            "ship": np.array(ship, dtype=np.float64),
            "acquisition": np.array(acq, dtype=np.float64),
            "actual order": np.array(order, dtype=np.float64),
            "profit": np.array(profit, dtype=np.float64),
        }
        return self.observe(), reward, self.t >= self.T, info

    # ------------------------------------------------------------------
    def state_vector(self) -> Dict[str, np.ndarray]:
        """Persistent integer state in the product's SoA naming (for bit-exact comparisons)."""
        bt = [v for i in self.split_nodes for v in self.backlog_to[i]]
        return {
            "inv": np.array(self.inv, dtype=np.int32),
            "backlog": np.array(self.backlog, dtype=np.int32),
            "order_u": np.array(self.order_u, dtype=np.int32),
            "pipe": np.array([self.pipe[i][k] for i in range(self.m) for k in range(self.delay[i])], dtype=np.int32),
            "hist_d": np.array(self.hist_d, dtype=np.int32).reshape(-1),
            "hist_o": np.array(self.hist_o, dtype=np.int32).reshape(-1),
            "carry": np.array(self.carry, dtype=np.int32),
            "backlog_to": np.array(bt, dtype=np.int32),
        }


# --------------------------------------------------------------------------------------
# base-stock policy + DFO objective — base_restock_policy.py
# --------------------------------------------------------------------------------------
def base_stock_action(env: OracleEnv, z: Sequence[float]) -> np.ndarray:
    """base_stock_policy — base_restock_policy.py:4-21."""
    out = np.zeros(env.m)
    for i in range(env.m):
        inv_ech = float(env.inv[i]) + float(env.order_u[i]) - float(env.backlog[i])
        out[i] = min(float(env.order_max[i]), max(float(z[i]) - inv_ech, 0.0))
    return out


def base_stock_rollout(env: OracleEnv, z: Sequence[float], customer_demand, delay_mask=None):
    """The loop of dfo_func (base_restock_policy.py:30-39) / inv_management.py:219-231.
    Returns the per-period reward list (scalar for IM kinds, [m] arrays for MAIM kinds)."""
    env.reset(customer_demand, delay_mask)
    rewards = []
    done = False
    while not done:
        _, r, done, _ = env.step(base_stock_action(env, z))
        rewards.append(r)
    return rewards


def poisson_pmf(k: int, mu: float) -> float:
    """scipy.stats.poisson.pmf(k, mu) = exp(k*log(mu) - lgamma(k+1) - mu)  (scipy
    _discrete_distns.py poisson_gen._pmf via special.xlogy/gammaln)."""
    if k < 0:
        return 0.0
    return math.exp((k * math.log(mu) if k > 0 else 0.0) - math.lgamma(k + 1) - mu)


def dfo_value(env: OracleEnv, z: Sequence[float], customer_demand, pmf, delay_mask=None) -> float:
    """dfo_func — base_restock_policy.py:24-45: -1 / T * np.sum(prob * rewards).
    ``pmf`` is the probability array of the demand trace as the caller evaluates it: [T] for a serial env, [R, T]
    for a divergent one (``env.dist.pmf(env.customer_demand)``, :42), where the product broadcasts over the retailer
    rows and np.sum runs over the flattened [R, T] array (:45).  ``delay_mask`` replays noisy delays (the reference's
    flag is sticky, so dfo_func after a noisy reset rolls out with them)."""
    rewards = base_stock_rollout(env, z, customer_demand, delay_mask)
    prod = np.asarray(pmf, dtype=np.float64) * np.asarray(rewards, dtype=np.float64)
    want = -1 / env.T * np.sum(prod)
    assert want == -1 / env.T * np_sum_order(list(np.ascontiguousarray(prod).reshape(-1)))   # the restated order IS numpy's
    return want


# --------------------------------------------------------------------------------------
# centralised-critic observer — models/CC_Model.py:196-214 (+ FillInActions :165-193)
# --------------------------------------------------------------------------------------
def central_critic_flat(agent_obs: np.ndarray, actions=None, clip=(-1.0, 1.0)) -> np.ndarray:
    """agent_obs [m, O] → [m, (m-1) + (m-1)*O + O]: per agent the dict {own_obs, opponent_obs,
    opponent_action} of central_critic_observer flattened in sorted key order (opponent_action,
    opponent_obs, own_obs).  Opponents in agent order skipping the agent itself; opponent actions are
    the same step's actions clipped to [a, b] (CC_inv_management.py:523) or zeros (CC_Model.py:207)."""
    m, O = agent_obs.shape
    out = np.zeros((m, (m - 1) * (1 + O) + O))
    for i in range(m):
        others = [j for j in range(m) if j != i]
        for slot, j in enumerate(others):
            if actions is not None:
                out[i, slot] = min(max(float(actions[j]), clip[0]), clip[1])
            out[i, (m - 1) + slot * O:(m - 1) + (slot + 1) * O] = agent_obs[j]
        out[i, (m - 1) * (1 + O):] = agent_obs[i]
    return out


def eval_loop_accumulators(env: OracleEnv, customer_demand, actions, rescaled: bool = True, delay_mask=None):
    """The scripts' evaluation loop for one episode — MA_inv_management.py:538-587 (multi-agent kinds; same in
    CC_inv_management.py:512-556 and CC_inv_management_div.py:500-544), inv_management.py:570-606 (single-agent
    kinds) and, with ``rescaled=False``, the LP replay loops (DSHLP_4.py:896-928: ``sum(s[:, 0])`` on raw state).
    Returns ``[episode_reward, total_inventory, total_backlog, customer_backlog, stage_profit[0..m-1]]``."""
    m, T = env.m, env.T
    env.reset(np.array(customer_demand), delay_mask)
    a, b = float(env.a), float(env.b)
    episode_reward = 0
    total_inventory = 0
    total_backlog = 0
    customer_backlog = 0
    stage_rewards = [0.0] * m
    for t in range(T):
        obs, reward, done, info = env.step(actions[t])
        undo = (lambda x, i: rev_scale(float(x), float(env.inv_max[i]), a, b)) if rescaled else (lambda x, i: float(x))
        if env.multi:                                   # MA_inv_management.py:568-581
            total_step_inv = 0
            total_step_bl = 0
            for i in range(m):
                episode_reward += reward[i]
                stage_rewards[i] += info["profit"][i]
                total_step_inv += undo(obs[i][0], i)
                total_step_bl += undo(obs[i][1], i)
            total_inventory += total_step_inv
            total_backlog += total_step_bl
            customer_backlog += undo(obs[0][1], 0)
        else:                                           # inv_management.py:585-598
            inv = [undo(obs[i][0], i) for i in range(m)]
            bl = [undo(obs[i][1], i) for i in range(m)]
            total_inventory += sum(inv)
            total_backlog += sum(bl)
            customer_backlog += bl[0]
            for i in range(m):
                stage_rewards[i] += info["profit"][i]
            episode_reward += reward
    return np.array([episode_reward, total_inventory, total_backlog, customer_backlog] + stage_rewards, dtype=np.float64)
