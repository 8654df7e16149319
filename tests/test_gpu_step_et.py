"""Env-per-thread step kernels (IMX_STEP_ET: one thread of the tile's compute group simulates a whole env; csrc/imx_step_tma.cuh
tile_period_et) against the C oracle on every env: one tile per CTA, the pipelined kernel and the multi-period kernel; the shipped
divergent networks, random trees up to 8 nodes, serial chains (forced), float32 observations and obs = NULL."""
import numpy as np
import pytest
import torch

from harness import random_tree_config
from marl_for_im_b200 import presets
from marl_for_im_b200.envs import ENV_CLASSES
from oracle import c_oracle

pytestmark = pytest.mark.gpu


def _cases():
    rng = np.random.default_rng(17)
    out = [("MAIM_div", presets.div1()), ("MAIM_div", presets.div2(share_network=True, prev_actions=True)), ("IM_div", presets.div2(prev_length=2)),
           ("IM_div", presets.div1(time_dependency=False, prev_demand=False)), ("MAIM", presets.serial4()), ("IM", presets.serial8(prev_actions=True)),
           ("MAIM", presets.serial2())]
    for trial in range(4):
        m = int(rng.integers(3, 9))
        cfg = random_tree_config(rng, m, int(rng.integers(2, 5)), periods=12, prev_actions=bool(trial % 2), prev_length=1 + trial % 2,
                                 independent=bool(trial == 1), share_network=bool(trial == 3))
        cfg["delay"] = np.minimum(cfg["delay"], 4)
        out.append(("MAIM_div" if trial % 2 else "IM_div", cfg))
    return out


CASES = _cases()


@pytest.mark.parametrize("case", range(len(CASES)))
@pytest.mark.parametrize("pipe,threads", [("0", "64"), ("1", "64"), ("1", "32"), ("0", "128")])
def test_step_et_matches_c_oracle(case, pipe, threads, monkeypatch):
    kind, cfg = CASES[case]
    monkeypatch.setenv("IMX_STEP_ET", "1")
    monkeypatch.setenv("IMX_STEP_ET_THREADS", threads)
    monkeypatch.setenv("IMX_PIPE", pipe)
    monkeypatch.setenv("IMX_PIPE_CTAS", "1")
    N = 2048 + 128 * case + (4 if case % 2 else 0)
    obs_dtype = "float32" if case % 4 == 2 else "float64"
    env = ENV_CLASSES[kind](dict(cfg, num_envs=N, obs_dtype=obs_dtype))
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
    obs_many, rew_many, done = env.step_many(a_dev[half:])
    assert (3 if pipe == "1" else 2) in variants
    want = c_oracle.COracle(kind, cfg).run(demand, actions)
    assert want["bad"] == 0 and int(env.error_flags.abs().sum()) == 0
    got_rew = torch.cat([torch.stack(rews), rew_many.reshape(T - half, N, -1)]).cpu().numpy()
    np.testing.assert_array_equal(got_rew, want["reward"])
    np_dt = np.float32 if obs_dtype == "float32" else np.float64
    np.testing.assert_array_equal(obs_many[-1].cpu().numpy(), want["obs_last"].astype(np_dt))
    st = {k: v.cpu().numpy() for k, v in env.state_dict().items()}
    for k in ("inv", "backlog", "order_u", "pipe") + (("backlog_to",) if "backlog_to" in st else ()):
        np.testing.assert_array_equal(st[k], want[k], err_msg=k)
    if "hist_o" in st:
        np.testing.assert_array_equal(st["hist_o"], want["hist_o"].reshape(N, -1))
