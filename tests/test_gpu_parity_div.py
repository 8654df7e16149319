"""CUDA path vs the oracle / golden fixtures — divergent networks (pass-exact split, signed ledger)."""
import itertools

import numpy as np
import pytest

from cuda_harness import assert_same, run_cuda
from harness import golden_names, load_golden, make_delay_mask, random_case, run_oracle
from marl_for_im_b200 import presets

pytestmark = pytest.mark.gpu

DIV_GOLDEN = [n for n in golden_names() if "div" in n]
MODES = list(itertools.product([False, True], repeat=3))


@pytest.mark.parametrize("name", DIV_GOLDEN)
def test_cuda_matches_golden_div(name):
    g = load_golden(name)
    got = run_cuda(g["kind"], g["config"], g["demand_trace"], g["actions"], g["delay_mask"], n_copies=72)
    assert_same(g["ref"], got, name)


@pytest.mark.parametrize("kind,preset", [("MAIM_div", "div1"), ("IM_div", "div1"), ("MAIM_div", "div2"), ("IM_div", "div2")])
def test_cuda_matches_oracle_div_modes(kind, preset):
    rng = np.random.default_rng(123)
    for td, pd, pa in MODES:
        if kind == "MAIM_div" and (not td) and pa and (not pd):
            continue
        for P, mu, amode, indep in [(1, 5, "uniform", False), (2, 9, "near_eq", True)]:
            cfg = presets.PRESETS[preset](time_dependency=td, prev_demand=pd, prev_actions=pa, prev_length=P, independent=indep)
            m = cfg["num_nodes"]
            cfg["inv_max"] = np.array([30, 25, 40, 35, 30, 45][:m], dtype=float)
            cfg["inv_target"] = np.array([0, 3, 5.5, 1, 0, 2][:m], dtype=float)
            if kind == "MAIM_div":
                cfg["share_network"] = bool(P == 2)
            demand, actions = random_case(kind, cfg, rng, mu=mu, action_mode=amode)
            assert_same(run_oracle(kind, cfg, demand, actions), run_cuda(kind, cfg, demand, actions, n_copies=72),
                        f"{kind}/{preset}/{(td, pd, pa)}/P{P}")


@pytest.mark.parametrize("kind", ["MAIM_div", "IM_div"])
def test_cuda_div_noisy_and_raw(kind):
    rng = np.random.default_rng(8)
    cfg = presets.div2()
    for _ in range(4):
        demand, actions = random_case(kind, cfg, rng, mu=7, action_mode="near_eq")
        mask = make_delay_mask(kind, cfg["delay"], 30, 0.3, rng)
        assert_same(run_oracle(kind, cfg, demand, actions, mask), run_cuda(kind, cfg, demand, actions, mask, n_copies=72), "noisy")
    if kind == "IM_div":
        cfg = presets.div1(prev_actions=True)
        cfg["standardise_state"] = False
        cfg["standardise_actions"] = False
        demand, actions = random_case(kind, cfg, rng, mu=12)
        assert_same(run_oracle(kind, cfg, demand, actions), run_cuda(kind, cfg, demand, actions, n_copies=72), "raw")


def test_cuda_wide_tree():
    """A bushier tree: 11 nodes, one node with 4 children, depth 3, heterogeneous lead times."""
    rng = np.random.default_rng(31)
    conn = {0: [1, 2], 1: [3, 4, 5, 6], 2: [7], 3: [], 4: [8, 9, 10], 5: [], 6: [], 7: [], 8: [], 9: [], 10: []}
    m = 11
    cfg = {"num_nodes": m, "num_periods": 20, "connections": conn, "init_inv": np.ones(m) * 12, "inv_target": np.ones(m) * 1,
           "inv_max": rng.integers(20, 45, m).astype(float), "stock_cost": rng.uniform(0.1, 0.5, m),
           "backlog_cost": rng.uniform(0.3, 0.9, m), "delay": rng.integers(1, 4, m), "time_dependency": True,
           "prev_demand": True, "prev_actions": True, "prev_length": 2, "independent": False, "share_network": True}
    for kind in ("MAIM_div", "IM_div"):
        for amode in ("uniform", "near_eq"):
            demand, actions = random_case(kind, cfg, rng, mu=4, action_mode=amode)
            assert_same(run_oracle(kind, cfg, demand, actions), run_cuda(kind, cfg, demand, actions, n_copies=72), f"tree {kind}")


def test_cuda_div_batch_distinct_envs():
    import torch
    from marl_for_im_b200.envs import MultiAgentInvManagementDiv
    cfg = presets.div2()
    N, T, m, R = 2048 + 8, 30, 6, 3
    rng = np.random.default_rng(420)
    demand = rng.poisson(5, size=(N, R, T)).astype(np.int32)
    actions = np.clip(rng.normal(-0.6, 0.5, size=(T, N, m)), -1, 1)
    env = MultiAgentInvManagementDiv(dict(cfg, num_envs=N))
    env.reset(customer_demand=demand)
    a_dev = torch.as_tensor(actions, device="cuda:0")
    obs, rew = [], []
    for t in range(T):
        o, r, done, _ = env.step(a_dev[t])
        obs.append(torch.stack([o[n] for n in env.agent_names], dim=1).cpu().numpy())
        rew.append(torch.stack([r[n] for n in env.agent_names], dim=1).cpu().numpy())
    obs, rew = np.stack(obs), np.stack(rew)
    assert int(env.error_flags.abs().sum()) == 0
    for n in list(range(0, N, 61)) + [N - 1]:
        want = run_oracle("MAIM_div", cfg, demand[n], actions[:, n])
        np.testing.assert_array_equal(want["obs"][1:], obs[:, n])
        np.testing.assert_array_equal(want["reward"], rew[:, n])


def test_device_generated_noisy_delay_mask():
    """Batched envs draw the noisy-delay outcomes from Philox on the device: the drawn mask is exposed,
    its rate matches the threshold, and the dynamics under it match the oracle replaying that mask."""
    import torch
    from marl_for_im_b200.envs import MultiAgentInvManagement, MultiAgentInvManagementDiv
    for cls, kind, cfg in ((MultiAgentInvManagement, "MAIM", presets.serial4()), (MultiAgentInvManagementDiv, "MAIM_div", presets.div2())):
        N, T = 4096, 30
        m = cfg.get("num_nodes", cfg.get("num_stages"))
        rng = np.random.default_rng(44)
        R = 3 if kind == "MAIM_div" else 1
        demand = rng.poisson(5, size=(N, R, T)).astype(np.int32)
        actions = np.clip(rng.normal(-0.5, 0.5, size=(T, N, m)), -1, 1)
        env = cls(dict(cfg, num_envs=N, seed=5))
        env.reset(customer_demand=demand if R > 1 else demand[:, 0], noisy_delay=True, noisy_delay_threshold=0.3)
        mask = env.delay_mask_device().cpu().numpy()                       # [T, N, m]
        assert abs(mask.mean() - 0.3) < 0.01
        a_dev = torch.as_tensor(actions, device="cuda:0")
        for t in range(T):
            o, r, _, _ = env.step(a_dev[t])
        obs = torch.stack([o[n] for n in env.agent_names], dim=1).cpu().numpy()
        for n in (0, 77, 4095):
            want = run_oracle(kind, cfg, demand[n] if R > 1 else demand[n, 0], actions[:, n], mask[:, n].astype(bool))
            np.testing.assert_array_equal(obs[n], want["obs"][-1])
        # a different threshold re-creates the generator state (handle rebuilt) and changes the rate
        env.reset(customer_demand=demand if R > 1 else demand[:, 0], noisy_delay=True, noisy_delay_threshold=0.6)
        assert abs(env.delay_mask_device().float().mean().item() - 0.6) < 0.01


@pytest.mark.parametrize("kind", ["MAIM_div", "IM_div"])
def test_cuda_random_divergent_networks(kind):
    """Random trees (5-12 nodes, up to 5 children per node) against the oracle (itself pinned to the reference on the
    same generator by tests/test_oracle_vs_reference.py and the *div_tree* fixtures)."""
    from harness import random_tree_config
    rng = np.random.default_rng(3027 if kind == "MAIM_div" else 3028)
    for trial in range(10):
        m = int(rng.integers(5, 13))
        cfg = random_tree_config(rng, m, int(rng.integers(2, 6)), periods=16, prev_actions=bool(trial % 2), prev_length=1 + trial % 3,
                                 independent=bool(trial % 3 == 0), share_network=(kind == "MAIM_div" and trial % 4 == 0))
        demand, actions = random_case(kind, cfg, rng, mu=4, action_mode="near_eq" if trial % 2 else "uniform")
        assert_same(run_oracle(kind, cfg, demand, actions), run_cuda(kind, cfg, demand, actions, n_copies=40), f"{kind} tree {cfg['connections']}")


@pytest.mark.parametrize("obs_dtype", ["float64", "float32"])
def test_rotated_observation_rows_wide(obs_dtype, monkeypatch):
    """Observation rows of 16 elements (128 bytes in float64: eight 16-byte chunks, rotation over all 8 lanes of a
    quarter-warp; 64 bytes in float32: four chunks) through the runtime-specialised kernels, plain and multi-period."""
    import torch
    from harness import copy_config, random_tree_config
    from marl_for_im_b200.envs import ENV_CLASSES
    monkeypatch.setenv("IMX_JIT", "1")
    rng = np.random.default_rng(16)
    cfg = random_tree_config(rng, 7, 3, periods=12, prev_actions=True, prev_length=4, share_network=True)
    cfg["delay"] = np.array([4, 1, 2, 4, 3, 1, 2])
    kind, n = "MAIM_div", 256
    env = ENV_CLASSES[kind](dict(copy_config(cfg), num_envs=n, obs_dtype=obs_dtype))
    assert env.obs_len == 16
    demand, actions = random_case(kind, cfg, rng, mu=4, action_mode="near_eq")
    want = run_oracle(kind, cfg, demand, actions)
    d = np.broadcast_to(np.asarray(demand)[None], (n,) + np.asarray(demand).shape)
    a = torch.as_tensor(np.broadcast_to(actions[:, None, :], (actions.shape[0], n, 7)).copy(), device="cuda:0")
    cast = (lambda x: x.astype(np.float32)) if obs_dtype == "float32" else (lambda x: x)
    env.reset(customer_demand=d)
    for t in range(12):
        env.step(a[t])
        assert env._lib.imx_kernel_variant(env._handle) in (2, 3)
        got = env.last_obs.cpu().numpy()
        for k in (0, 3, 8, 255):
            np.testing.assert_array_equal(got[k], cast(want["obs"][t + 1]), err_msg=f"t={t} env={k}")
    env.reset(customer_demand=d)
    o, r, _ = env.step_many(a)
    np.testing.assert_array_equal(o[:, 5].cpu().numpy(), cast(want["obs"][1:]))
    np.testing.assert_array_equal(r[:, 5].cpu().numpy(), want["reward"])
