"""The plain-C oracle (oracle/imx_oracle.c) against the reference-generated golden fixtures and the
Python oracle — so that it can be trusted for the full-size GPU checks."""
import numpy as np
import pytest

from harness import golden_names, load_golden, make_delay_mask, random_case, run_oracle
from marl_for_im_b200 import presets
from oracle import c_oracle, im_oracle


@pytest.mark.parametrize("name", golden_names())
def test_c_oracle_matches_golden(name):
    g = load_golden(name)
    co = c_oracle.COracle(g["kind"], g["config"])
    d = np.asarray(g["demand_trace"])[None]
    a = np.asarray(g["actions"])[:, None, :]
    mask = None if g["delay_mask"] is None else np.asarray(g["delay_mask"])[None]
    out = co.run(d, a, mask, all_obs=True)
    assert out["bad"] == 0
    np.testing.assert_array_equal(out["obs_all"][:, 0], g["ref"]["obs"])
    want_r = g["ref"]["reward"] if co.env.multi else g["ref"]["reward"][:, :1]
    np.testing.assert_array_equal(out["reward"][:, 0], want_r)
    np.testing.assert_array_equal(out["inv"][0], g["ref"]["inv"][-1])
    np.testing.assert_array_equal(out["backlog"][0], g["ref"]["backlog"][-1])
    np.testing.assert_array_equal(out["order_u"][0], g["ref"]["order_u"][-1])


def test_c_oracle_matches_python_oracle_batch():
    rng = np.random.default_rng(12)
    for kind, cfg in (("MAIM", presets.serial8(prev_actions=True, prev_length=3)), ("IM", presets.serial4_dfo()),
                      ("MAIM_div", presets.div2(share_network=True)), ("IM_div", presets.div1(prev_actions=True))):
        N = 40
        co = c_oracle.COracle(kind, cfg)
        dem, act, masks = [], [], []
        for n in range(N):
            d, a = random_case(kind, cfg, rng, mu=7, action_mode="near_eq" if n % 2 else "uniform")
            dem.append(d)
            act.append(a)
            masks.append(make_delay_mask(kind, cfg["delay"], 30, 0.25, rng))
        out = co.run(np.stack(dem), np.stack(act, axis=1), np.stack(masks), all_obs=True, threads=4)
        for n in range(N):
            want = run_oracle(kind, cfg, dem[n], act[n], masks[n])
            np.testing.assert_array_equal(out["obs_all"][:, n], want["obs"])
            np.testing.assert_array_equal(out["reward"][:, n], want["reward"] if co.env.multi else want["reward"][:, :1])
            np.testing.assert_array_equal(out["inv"][n], want["inv"][-1])


def test_c_oracle_rollout_matches_python():
    from scipy.stats import poisson
    rng = np.random.default_rng(3)
    for kind, preset in (("IM", "serial4_dfo"), ("MAIM", "serial8")):
        cfg = presets.PRESETS[preset]()
        cfg.update(time_dependency=False, prev_demand=False, prev_actions=False, standardise_state=False, standardise_actions=False)
        m = cfg["num_stages"]
        N = 30
        demand = rng.poisson(5, size=(N, 30))
        z = rng.integers(5, 40, size=(N, m)) + rng.choice([0.0, 0.37], size=(N, m))
        co = c_oracle.COracle(kind, cfg)
        out = co.rollout(z, demand, pmf=poisson.pmf(demand, mu=5), step_rewards=True)
        for n in range(N):
            env = im_oracle.OracleEnv(kind, cfg)
            rewards = im_oracle.base_stock_rollout(env, z[n], demand[n])
            np.testing.assert_array_equal(out["step_rewards"][:, n], np.array(rewards).reshape(30, -1))
            np.testing.assert_array_equal(out["inv"][n], np.array(env.inv))
            if kind == "IM":
                assert out["dfo"][n] == im_oracle.dfo_value(env, z[n], demand[n], poisson.pmf(demand[n], mu=5))


def test_full_size_properties_c_oracle():
    """Size-independent invariants at BASELINE size (65 536 envs x 30 periods) on the CPU checker:
    bounds 0 <= inv, order_u <= inv_max, 0 <= backlog <= inv_max, observations inside [a, b] for
    the clipped fields, and goods conservation of a serial chain without noisy delay."""
    cfg = presets.serial4()
    N, T, m = 65536, 30, 4
    rng = np.random.default_rng(420)
    demand = rng.poisson(5, size=(N, T)).astype(np.int32)
    actions = np.random.default_rng(0).uniform(-1, 1, size=(T, N, m))
    out = c_oracle.COracle("MAIM", cfg).run(demand, actions)
    for k in ("inv", "order_u", "backlog"):
        assert out[k].min() >= 0 and out[k].max() <= 30
    assert np.all(out["obs_last"][:, :, :3] >= -1) and np.all(out["obs_last"][:, :, :3] <= 1)
    assert np.isfinite(out["reward"]).all()
