"""dfo_func on divergent envs (inv_management_div.py:228) and under a sticky noisy delay (MAIM_env.py:192-194,449-457):
the CUDA rollout + objective kernels against reference-generated values (tests/golden/dfo/) and the oracle."""
import numpy as np
import pytest
import torch
from scipy.stats import poisson

from dfo_cases import load_dfo_cases
from marl_for_im_b200 import presets
from marl_for_im_b200.base_restock_policy import base_stock_policy, dfo_func, dfo_func_batch
from marl_for_im_b200.envs import ENV_CLASSES
from oracle import im_oracle

pytestmark = pytest.mark.gpu
CASES = load_dfo_cases()


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_dfo_matches_reference_values(case):
    kind, cfg = case["kind"], case["config"]
    K = case["z"].shape[0]
    div = kind.endswith("_div")
    dem = case["demand"] if div else case["demand"][:, 0]
    pmf = poisson.pmf(dem, mu=5)
    # batched: one env per stored call
    env = ENV_CLASSES[kind](dict(cfg, num_envs=K, noisy_delay=case["noisy"]) if div else dict(cfg, num_envs=K))
    out = env.rollout_basestock(case["z"], customer_demand=dem, pmf=pmf, delay_mask=case["mask"] if case["noisy"] else None)
    np.testing.assert_array_equal(out["dfo"].cpu().numpy(), case["dfo"])
    # drop-in (N = 1): the reference's call signature; the sticky flag is set by a noisy reset, the rollout's uniforms
    # come from numpy's global stream in the reference's order -> replay them through a patched np.random.uniform
    env1 = ENV_CLASSES[kind](dict(cfg))
    for k in range(K):
        saved = np.random.uniform
        try:
            if case["noisy"]:
                from harness import _uniform_replayer
                env1.reset(customer_demand=dem[k], noisy_delay=True, noisy_delay_threshold=0.5)
                seq = iter(_uniform_replayer(kind, [int(d) for d in env1.delay], env1.num_periods, case["mask"][k]))
                np.random.uniform = lambda *a, **kw: 0.0 if next(seq) == 0.0 else 1.0
            got = dfo_func(case["z"][k], env1, dem[k])
        finally:
            np.random.uniform = saved
        assert got == case["dfo"][k], f"{case['name']}[{k}] drop-in"


@pytest.mark.parametrize("R_preset", ["chain", "div1", "div2"])
def test_dfo_div_batch_vs_oracle(R_preset):
    """R = 1, 2, 3 retailer rows; every env of a 257-env batch against the oracle's dfo_value."""
    if R_preset == "chain":
        cfg = {"num_nodes": 4, "connections": {0: [1], 1: [2], 2: [3], 3: []}, "num_periods": 30, "init_inv": np.ones(4) * 10,
               "stock_cost": np.array([0.35, 0.3, 0.4, 0.2]), "backlog_cost": np.array([0.5, 0.7, 0.6, 0.9]), "demand_dist": "poisson",
               "inv_target": np.zeros(4), "inv_max": np.ones(4) * 30, "delay": np.array([1, 2, 3, 1]), "mu": 5}
    else:
        cfg = presets.PRESETS[R_preset]()
        cfg.update(demand_dist="poisson", mu=5)
    cfg.update(time_dependency=False, prev_demand=False, prev_actions=False, standardise_state=False, standardise_actions=False)
    rng = np.random.default_rng(31)
    N, T, m = 257, 30, cfg["num_nodes"]
    env = ENV_CLASSES["IM_div"](dict(cfg, num_envs=N))
    R = len(env._retailers)
    assert R == {"chain": 1, "div1": 2, "div2": 3}[R_preset]
    demand = rng.poisson(5, size=(N, R, T))
    z = rng.integers(5, 41, size=(N, m)).astype(np.float64) + rng.choice([0.0, 0.37], size=(N, m))
    pmf = poisson.pmf(demand, mu=5)
    got = env.rollout_basestock(z, customer_demand=demand, pmf=pmf)["dfo"].cpu().numpy()
    orc = im_oracle.OracleEnv("IM_div", dict(cfg))
    for n in range(0, N, 7):
        assert got[n] == im_oracle.dfo_value(orc, z[n], demand[n], pmf[n]), n
    # population form: K policies x D traces in one launch
    K, D = 5, 11
    envb = ENV_CLASSES["IM_div"](dict(cfg, num_envs=K * D))
    gb = dfo_func_batch(z[:K], envb, demand[:D]).cpu().numpy()
    for k in range(K):
        for d in range(0, D, 3):
            assert gb[k, d] == im_oracle.dfo_value(orc, z[k], demand[d], pmf[d])


@pytest.mark.parametrize("kind,preset", [("MAIM", "serial4"), ("IM", "serial8"), ("MAIM_div", "div2"), ("IM_div", "div1")])
def test_noisy_rollout_equals_step_loop(kind, preset):
    """Fused rollout under noisy delays == reset(noisy) + T x step(base_stock_policy) on the same replayed mask, and on the
    Philox mask the env itself generated (counter-based: the rollout re-derives the same draws)."""
    cfg = presets.PRESETS[preset]()
    cfg.update(time_dependency=False, prev_demand=False, prev_actions=False)
    if kind in ("IM", "MAIM", "IM_div"):
        cfg.update(standardise_state=False, standardise_actions=False)
    N, T = 193, cfg["num_periods"]
    env = ENV_CLASSES[kind](dict(cfg, num_envs=N))
    m, R = env.num_nodes, len(env._retailers)
    rng = np.random.default_rng(8)
    demand = rng.poisson(6, size=(N, R, T)).astype(np.int32)
    z = rng.integers(5, 41, size=(N, m)).astype(np.float64)
    if kind == "MAIM_div":
        z = rng.uniform(-1, 1, size=(N, m))
    mask = rng.uniform(size=(N, T, m)) <= 0.3
    # (a) replayed mask: oracle + step loop
    out = env.rollout_basestock(z, customer_demand=demand, step_rewards=True, write_state=True, delay_mask=mask)
    sr = out["step_rewards"].cpu().numpy()
    st = {k: v.cpu().numpy().copy() for k, v in env.state_dict().items()}
    for n in range(0, N, 17):
        orc = im_oracle.OracleEnv(kind, dict(cfg))
        rewards = im_oracle.base_stock_rollout(orc, z[n], demand[n] if env.DIV else demand[n, 0], mask[n])
        np.testing.assert_array_equal(sr[:, n], np.array(rewards).reshape(sr[:, n].shape))
        want = orc.state_vector()
        for k in ("inv", "backlog", "order_u", "pipe", "carry"):
            if k in want and k in st:
                np.testing.assert_array_equal(st[k][n], want[k], err_msg=k)
    env.reset(customer_demand=demand, delay_mask=mask)
    zt = torch.as_tensor(z, device="cuda:0")
    for t in range(T):
        _, r, done, _ = env.step(base_stock_policy(zt, env))
        rt = torch.stack([r[a] for a in env.agent_names], dim=1).cpu().numpy() if env.MULTI else r.cpu().numpy()
        np.testing.assert_array_equal(rt, sr[t])
    for k, v in env.state_dict().items():
        np.testing.assert_array_equal(v.cpu().numpy(), st[k], err_msg=k)
    # (b) Philox mask: reset(noisy) draws it, the rollout of the same episode id re-derives it
    env2 = ENV_CLASSES[kind](dict(cfg, num_envs=N, noisy_delay=True, noisy_delay_threshold=0.35) if env.DIV else dict(cfg, num_envs=N))
    env2.reset(customer_demand=demand, noisy_delay=True, noisy_delay_threshold=0.35)
    ep = env2._episode
    total = None
    for t in range(T):
        _, r, done, _ = env2.step(base_stock_policy(zt, env2))
        rt = torch.stack([r[a] for a in env2.agent_names], dim=1) if env2.MULTI else r
        total = rt.clone() if total is None else total + rt
    env2._episode = ep - 1                                  # rollout_basestock increments: same episode id -> same draws
    fused = env2.rollout_basestock(z, customer_demand=demand)["returns"]
    assert torch.equal(fused, total)
