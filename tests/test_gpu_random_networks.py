"""Specialised kernels on networks beyond the shipped presets: step, step_many, the pipelined step and the fused rollout against
the C oracle on every env — random trees with up to 12 nodes and 5 children per node (lane mapping, more than 8 nodes: the lanes
rollout), serial chains whose length is not a power of two, and lanes == env-per-thread rollouts on Philox demand."""
import numpy as np
import pytest
import torch

from harness import random_tree_config
from marl_for_im_b200 import presets
from marl_for_im_b200.envs import ENV_CLASSES
from oracle import c_oracle

pytestmark = pytest.mark.gpu


def _cfgs():
    rng = np.random.default_rng(99)
    out = [("MAIM_div", presets.div1()), ("MAIM_div", presets.div2(share_network=True, prev_actions=True)), ("IM_div", presets.div2(prev_length=2)),
           ("IM_div", presets.div1(time_dependency=False, prev_demand=False))]
    for trial in range(4):
        m = int(rng.integers(5, 13))
        kind = "MAIM_div" if trial % 2 else "IM_div"
        out.append((kind, random_tree_config(rng, m, int(rng.integers(2, 6)), periods=12, prev_actions=bool(trial % 2), prev_length=1 + trial % 3,
                                             independent=bool(trial == 1), share_network=(kind == "MAIM_div" and trial == 3))))
    return out


CFGS = _cfgs()


@pytest.mark.parametrize("case", range(len(CFGS)))
@pytest.mark.parametrize("pipe", ["0", "1"])
def test_step_on_random_networks_matches_c_oracle(case, pipe, monkeypatch):
    kind, cfg = CFGS[case]
    monkeypatch.setenv("IMX_PIPE", pipe)
    monkeypatch.setenv("IMX_PIPE_CTAS", "1")
    N = 2048 + 64 * case + (4 if case % 2 else 0)            # whole tiles (32 envs) plus, for odd cases, a tail for the direct kernel
    env = ENV_CLASSES[kind](dict(cfg, num_envs=N))
    m, T, R = env.num_nodes, env.num_periods, len(env._retailers)
    rng = np.random.default_rng(case)
    demand = rng.poisson(5, size=(N, R, T)).astype(np.int32)
    actions = np.clip(rng.normal(-0.5, 0.5, size=(T, N, m)), -1.1, 1.1)
    env.reset(customer_demand=demand)
    a_dev = torch.as_tensor(actions, device="cuda:0")
    rews, variants = [], set()
    half = T // 2
    for t in range(half):
        o, r, done, _ = env.step(a_dev[t])
        variants.add(env._lib.imx_kernel_variant(env._handle))
        rews.append(torch.stack([r[n] for n in env.agent_names], dim=1) if env.MULTI else r[:, None])
    obs_many, rew_many, done = env.step_many(a_dev[half:])                     # the rest through the multi-period kernel
    assert (3 if pipe == "1" else 2) in variants
    want = c_oracle.COracle(kind, cfg).run(demand, actions)
    assert want["bad"] == 0 and int(env.error_flags.abs().sum()) == 0
    got_rew = torch.cat([torch.stack(rews), rew_many.reshape(T - half, N, -1)]).cpu().numpy()
    np.testing.assert_array_equal(got_rew, want["reward"])
    np.testing.assert_array_equal(obs_many[-1].cpu().numpy(), want["obs_last"])
    st = {k: v.cpu().numpy() for k, v in env.state_dict().items()}
    for k in ("inv", "backlog", "order_u", "pipe", "backlog_to"):
        np.testing.assert_array_equal(st[k], want[k], err_msg=k)


@pytest.mark.parametrize("case", range(len(CFGS)))
def test_rollout_on_random_networks_matches_c_oracle(case, monkeypatch):
    kind, cfg = CFGS[case]
    cfg = dict(cfg, time_dependency=False, prev_demand=False, prev_actions=False)
    if kind == "IM_div":
        cfg.update(standardise_state=False, standardise_actions=False)
    N = 2048 + 32 * case + (7 if case % 2 else 0)
    env = ENV_CLASSES[kind](dict(cfg, num_envs=N))
    m, T, R = env.num_nodes, env.num_periods, len(env._retailers)
    rng = np.random.default_rng(100 + case)
    demand = rng.poisson(6, size=(N, R, T)).astype(np.int32)
    z = rng.integers(5, 41, size=(N, m)).astype(np.float64) if kind == "IM_div" else rng.uniform(-1, 1, size=(N, m))
    mask = (rng.uniform(size=(N, T, m)) <= 0.3) if case % 3 == 0 else None
    out = env.rollout_basestock(z, customer_demand=demand, step_rewards=True, write_state=True, delay_mask=mask)
    assert env._lib.imx_kernel_variant(env._handle) == 2
    want = c_oracle.COracle(kind, cfg).rollout(z, demand, step_rewards=True, delay_mask=mask)
    assert want["bad"] == 0 and int(env.error_flags.abs().sum()) == 0
    np.testing.assert_array_equal(out["step_rewards"].cpu().numpy().reshape(T, N, -1), want["step_rewards"])
    np.testing.assert_array_equal(out["returns"].cpu().numpy().reshape(N, -1), want["returns"])
    st = {k: v.cpu().numpy() for k, v in env.state_dict().items()}
    for k in ("inv", "backlog", "order_u", "pipe", "backlog_to"):
        np.testing.assert_array_equal(st[k], want[k], err_msg=k)


@pytest.mark.parametrize("kind,cfg", [("MAIM", dict(presets.serial4(), num_stages=3, init_inv=np.ones(3) * 10, inv_target=np.zeros(3), inv_max=np.ones(3) * 30,
                                                    price=np.array([4, 3, 2, 1]), stock_cost=np.array([0.35, 0.3, 0.4]),
                                                    backlog_cost=np.array([0.5, 0.7, 0.6]), delay=np.array([1, 2, 3]))),
                                      ("IM", presets.serial8(prev_actions=True)), ("MAIM", presets.serial4(independent=True))])
def test_serial_chains_of_odd_length(kind, cfg, monkeypatch):
    N = 4096 + 32
    env = ENV_CLASSES[kind](dict(cfg, num_envs=N))
    m, T = env.num_nodes, env.num_periods
    rng = np.random.default_rng(5)
    demand = rng.poisson(5, size=(N, 1, T)).astype(np.int32)
    actions = rng.uniform(-1.1, 1.1, size=(T, N, m))
    env.reset(customer_demand=demand)
    a_dev = torch.as_tensor(actions, device="cuda:0")
    rews = []
    for t in range(T):
        o, r, done, _ = env.step(a_dev[t])
        rews.append(torch.stack([r[n] for n in env.agent_names], dim=1) if env.MULTI else r[:, None])
    want = c_oracle.COracle(kind, cfg).run(demand, actions)
    np.testing.assert_array_equal(torch.stack(rews).cpu().numpy(), want["reward"])
    obs_last = (torch.stack([o[n] for n in env.agent_names], dim=1) if env.MULTI else o).cpu().numpy()
    np.testing.assert_array_equal(obs_last, want["obs_last"])
    np.testing.assert_array_equal(env.state_dict()["pipe"].cpu().numpy(), want["pipe"])


def test_et_and_lanes_rollouts_agree_on_philox_demand(monkeypatch):
    cfg = presets.div2(time_dependency=False, prev_demand=False)
    cfg.update(demand_dist="poisson", mu=5)
    outs = []
    for et in ("0", "1"):
        monkeypatch.setenv("IMX_ROLLOUT_ET", et)
        env = ENV_CLASSES["MAIM_div"](dict(cfg, num_envs=5000, seed=7))
        env._episode = 41
        outs.append(env.rollout_basestock(np.full(6, 0.2))["returns"].clone())
    assert torch.equal(outs[0], outs[1])
