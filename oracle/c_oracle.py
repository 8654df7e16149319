"""TEST INFRASTRUCTURE — ctypes front end of oracle/imx_oracle.c (the plain-C restatement used for
full-size batches and as the strong CPU baseline).  Config derivation comes from
oracle/im_oracle.py (pinned to the reference); the C code only runs the dynamics.
Built by ``__graft_entry__.build()`` or on demand here (gcc -O2 -fopenmp -ffp-contract=off)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import im_oracle

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "imx_oracle.c")
LIB = os.path.join(HERE, "libimx_oracle.so")
MAXN, MAXC = 32, 8


class OrcCfg(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("multi", "div", "m", "T", "P", "D", "O", "R", "std_state", "std_actions", "cap_backlog",
                                         "independent", "share_network", "td", "pd", "pa", "wd_mult1", "wd_mult")] + \
               [("a", C.c_double), ("b", C.c_double)] + \
               [(k, C.c_int32 * MAXN) for k in ("inv_init", "inv_max", "order_max", "demand_max", "delay", "parent", "nchild", "retailer_idx")] + \
               [("children", (C.c_int32 * MAXC) * MAXN)] + \
               [(k, C.c_double * MAXN) for k in ("p", "c", "h", "bc", "target")]


_lib = None


def build(force=False):
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        subprocess.run(["gcc", "-O2", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off", "-o", LIB, SRC, "-lm"], check=True)
    return LIB


def load():
    global _lib
    if _lib is None:
        build()
        lib = C.CDLL(LIB)
        assert lib.orc_cfg_size() == C.sizeof(OrcCfg), "orc_cfg layout mismatch"
        lib.orc_run.restype = C.c_long
        lib.orc_rollout.restype = C.c_long
        _lib = lib
    return _lib


def max_threads():
    return int(load().orc_max_threads())


def make_cfg(env: im_oracle.OracleEnv) -> OrcCfg:
    c = OrcCfg()
    c.multi, c.div, c.m, c.T, c.P, c.D, c.O, c.R = int(env.multi), int(env.div), env.m, env.T, env.P, env.D, env.O, env.R
    c.std_state, c.std_actions = int(env.std_state), int(env.std_actions)
    c.cap_backlog = 1 if env.kind == "MAIM_div" else int(env.std_state)
    c.independent, c.share_network = int(env.independent), int(env.share_network)
    c.td, c.pd, c.pa = int(env.td), int(env.pd), int(env.pa)
    c.wd_mult1, c.wd_mult = (2, 1) if env.multi else (4, 2)
    c.a, c.b = env.a, env.b
    for i in range(env.m):
        c.inv_init[i], c.inv_max[i], c.order_max[i] = env.init_inv[i], env.inv_max[i], env.order_max[i]
        c.demand_max[i], c.delay[i] = env.demand_max[i], env.delay[i]
        c.p[i], c.c[i], c.h[i], c.bc[i], c.target[i] = env.p[i], env.c[i], env.stock_cost[i], env.backlog_cost[i], env.inv_target[i]
        c.retailer_idx[i] = -1
        c.parent[i] = env.parent[i] if env.div else -1
        ch = env.children[i] if env.div else []
        c.nchild[i] = len(ch)
        for k, v in enumerate(ch):
            c.children[i][k] = v
    for k, r in enumerate(env.retailers):
        c.retailer_idx[r] = k
    return c


def _p(a, ctype):
    return a.ctypes.data_as(C.POINTER(ctype)) if a is not None else None


class COracle:
    """Batched front end: the same trajectories as N independent im_oracle.OracleEnv instances."""

    def __init__(self, kind, config):
        self.env = im_oracle.OracleEnv(kind, dict(config))
        self.cfg = make_cfg(self.env)
        e = self.env
        self.m, self.T, self.O, self.R = e.m, e.T, e.O, e.R
        self.cols = e.m if e.multi else 1
        self.L = sum(e.delay)
        self.NB = sum(len(e.children[i]) for i in e.split_nodes) if e.div else 0

    def run(self, demand, actions, delay_mask=None, periods=None, all_obs=False, threads=0):
        """demand [N,R,T] (or [N,T]); actions [periods,N,m]; returns dict of arrays."""
        lib = load()
        N = demand.shape[0]
        periods = self.T if periods is None else periods
        demand = np.ascontiguousarray(demand.reshape(N, self.R, self.T), dtype=np.int32)
        actions = np.ascontiguousarray(actions, dtype=np.float64)
        assert actions.shape == (periods, N, self.m)
        mask = None if delay_mask is None else np.ascontiguousarray(delay_mask, dtype=np.uint8).reshape(N, self.T, self.m)
        out = {
            "obs_all": np.empty((periods + 1, N, self.m, self.O)) if all_obs else None,
            "obs_last": np.empty((N, self.m, self.O)),
            "reward": np.empty((periods, N, self.cols)),
            "inv": np.empty((N, self.m), dtype=np.int32), "backlog": np.empty((N, self.m), dtype=np.int32),
            "order_u": np.empty((N, self.m), dtype=np.int32), "pipe": np.empty((N, self.L), dtype=np.int32),
            "backlog_to": np.empty((N, max(self.NB, 1)), dtype=np.int32),
            "hist_d": np.empty((N, self.m, self.env.P), dtype=np.int32), "hist_o": np.empty((N, self.m, self.env.P), dtype=np.int32),
            "err": np.zeros(N, dtype=np.int32),
        }
        bad = lib.orc_run(C.byref(self.cfg), C.c_long(N), C.c_int(periods), _p(demand, C.c_int32), _p(actions, C.c_double),
                          _p(mask, C.c_ubyte), _p(out["obs_all"], C.c_double), _p(out["obs_last"], C.c_double),
                          _p(out["reward"], C.c_double), _p(out["inv"], C.c_int32), _p(out["backlog"], C.c_int32),
                          _p(out["order_u"], C.c_int32), _p(out["pipe"], C.c_int32), _p(out["backlog_to"], C.c_int32),
                          _p(out["hist_d"], C.c_int32), _p(out["hist_o"], C.c_int32), _p(out["err"], C.c_int32), C.c_int(threads))
        out["bad"] = int(bad)
        return out

    def rollout(self, z, demand, pmf=None, step_rewards=False, threads=0, delay_mask=None):
        lib = load()
        N = demand.shape[0]
        demand = np.ascontiguousarray(demand.reshape(N, self.R, self.T), dtype=np.int32)
        z = np.ascontiguousarray(z, dtype=np.float64)
        stride = 0 if z.size == self.m else self.m
        pmf = None if pmf is None else np.ascontiguousarray(np.broadcast_to(np.asarray(pmf, dtype=np.float64).reshape(-1, self.R, self.T), (N, self.R, self.T)))
        mask = None if delay_mask is None else np.ascontiguousarray(np.broadcast_to(np.asarray(delay_mask).reshape(-1, self.T, self.m), (N, self.T, self.m)), dtype=np.uint8)
        out = {"returns": np.empty((N, self.cols)), "step_rewards": np.empty((self.T, N, self.cols)) if step_rewards else None,
               "dfo": np.empty(N) if pmf is not None else None,
               "inv": np.empty((N, self.m), dtype=np.int32), "backlog": np.empty((N, self.m), dtype=np.int32),
               "order_u": np.empty((N, self.m), dtype=np.int32), "pipe": np.empty((N, self.L), dtype=np.int32),
               "backlog_to": np.empty((N, max(self.NB, 1)), dtype=np.int32)}
        bad = lib.orc_rollout(C.byref(self.cfg), C.c_long(N), _p(z, C.c_double), C.c_int(stride), _p(demand, C.c_int32),
                              _p(mask, C.c_ubyte), _p(pmf, C.c_double), _p(out["returns"], C.c_double), _p(out["step_rewards"], C.c_double),
                              _p(out["dfo"], C.c_double), _p(out["inv"], C.c_int32), _p(out["backlog"], C.c_int32),
                              _p(out["order_u"], C.c_int32), _p(out["pipe"], C.c_int32), _p(out["backlog_to"], C.c_int32), C.c_int(threads))
        out["bad"] = int(bad)
        return out
