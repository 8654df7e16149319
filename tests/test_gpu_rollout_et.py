"""The env-per-thread rollout kernel (csrc/imx_rollout_et.cuh: one thread simulates a whole env, network as compile-time lists)
against the C oracle on every env — all four kinds, the shipped networks, random trees up to 8 nodes, replayed noisy-delay masks,
final state — and against the lanes = nodes kernel on the Philox demand and Philox delay streams (same draws, same results)."""
import numpy as np
import pytest
import torch

from harness import random_tree_config
from marl_for_im_b200 import presets
from marl_for_im_b200.envs import ENV_CLASSES
from oracle import c_oracle

pytestmark = pytest.mark.gpu


def _cases():
    rng = np.random.default_rng(7)
    out = [("MAIM", presets.serial4()), ("MAIM", presets.serial8(independent=True)), ("IM", presets.serial8()), ("IM", presets.serial4_dfo()),
           ("MAIM", presets.serial2()), ("MAIM_div", presets.div1()), ("MAIM_div", presets.div2()), ("IM_div", presets.div2()), ("IM_div", presets.div1())]
    for trial in range(4):
        m = int(rng.integers(3, 9))
        out.append(("MAIM_div" if trial % 2 else "IM_div", random_tree_config(rng, m, int(rng.integers(2, 5)), periods=14, independent=bool(trial == 1))))
    return out


CASES = _cases()


@pytest.mark.parametrize("case", range(len(CASES)))
def test_et_rollout_matches_c_oracle(case, monkeypatch):
    kind, cfg = CASES[case]
    cfg = dict(cfg, time_dependency=False, prev_demand=False, prev_actions=False)
    if kind != "MAIM_div":
        cfg.update(standardise_state=False, standardise_actions=bool(case % 2 and kind == "MAIM"))
    if max(cfg["delay"]) > 4:
        cfg["delay"] = np.minimum(cfg["delay"], 4)
    monkeypatch.setenv("IMX_ROLLOUT_ET", "1")
    N = 2048 + 37 * case
    env = ENV_CLASSES[kind](dict(cfg, num_envs=N))
    m, T, R = env.num_nodes, env.num_periods, len(env._retailers)
    rng = np.random.default_rng(200 + case)
    demand = rng.poisson(6, size=(N, R, T)).astype(np.int32)
    scaled = kind == "MAIM_div" or cfg.get("standardise_actions", False)
    z = rng.uniform(-1, 1, size=(N, m)) if scaled else rng.integers(5, 41, size=(N, m)).astype(np.float64) + rng.choice([0.0, 0.37], size=(N, m))
    mask = (rng.uniform(size=(N, T, m)) <= 0.3) if case % 3 == 0 else None
    out = env.rollout_basestock(z, customer_demand=demand, step_rewards=True, write_state=True, delay_mask=mask)
    assert env._lib.imx_kernel_variant(env._handle) == 2
    want = c_oracle.COracle(kind, cfg).rollout(z, demand, step_rewards=True, delay_mask=mask)
    assert want["bad"] == 0 and int(env.error_flags.abs().sum()) == 0
    np.testing.assert_array_equal(out["step_rewards"].cpu().numpy().reshape(T, N, -1), want["step_rewards"])
    np.testing.assert_array_equal(out["returns"].cpu().numpy().reshape(N, -1), want["returns"])
    st = {k: v.cpu().numpy() for k, v in env.state_dict().items()}
    for k in ("inv", "backlog", "order_u", "pipe") + (("backlog_to",) if env.DIV and "backlog_to" in st else ()):
        np.testing.assert_array_equal(st[k], want[k], err_msg=k)
    # shared base-stock levels (z_stride = 0)
    out0 = env.rollout_basestock(z[0], customer_demand=demand)
    want0 = c_oracle.COracle(kind, cfg).rollout(z[0], demand)
    np.testing.assert_array_equal(out0["returns"].cpu().numpy().reshape(N, -1), want0["returns"])


@pytest.mark.parametrize("kind,preset", [("MAIM", "serial8"), ("MAIM_div", "div2"), ("IM", "serial4_dfo")])
def test_et_and_lanes_rollouts_agree_on_philox_streams(kind, preset, monkeypatch):
    cfg = presets.PRESETS[preset]()
    cfg.update(time_dependency=False, prev_demand=False, prev_actions=False, demand_dist="poisson", mu=5)
    outs = []
    for et in ("0", "1"):
        monkeypatch.setenv("IMX_ROLLOUT_ET", et)
        env = ENV_CLASSES[kind](dict(cfg, num_envs=4099, seed=11, noisy_delay=True, noisy_delay_threshold=0.25) if kind.endswith("div")
                                else dict(cfg, num_envs=4099, seed=11))
        m = env.num_nodes
        if not kind.endswith("div"):
            env.reset(noisy_delay=True, noisy_delay_threshold=0.25)     # sticky: the rollouts below run with Philox delays
        env._episode = 9
        z = np.full(m, 0.2) if kind == "MAIM_div" else np.full(m, 25.0)
        a = env.rollout_basestock(z, step_rewards=True)
        outs.append((a["returns"].clone(), a["step_rewards"].clone()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
