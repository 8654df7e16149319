"""Drop-in mode (no ``num_envs`` key): one environment, numpy / dict in and out, exactly the call
pattern of the reference scripts — checked against the fixtures generated from the reference."""
import numpy as np
import pytest

from harness import agent_names, copy_config, golden_names, load_golden
from marl_for_im_b200.envs import ENV_CLASSES

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", golden_names())
def test_dropin_episode_matches_reference_fixture(name):
    g = load_golden(name)
    kind, ref = g["kind"], g["ref"]
    multi = kind.startswith("MAIM")
    cfg = copy_config(g["config"])
    env = ENV_CLASSES[kind](cfg)
    if not kind.endswith("_div"):
        assert "num_stages" not in cfg and "inv_max" not in cfg          # quirk 11: serial ctors pop the caller's dict
    m, T = env.num_nodes, env.num_periods
    names = agent_names(kind, m)
    kw = {}
    if g["delay_mask"] is not None:
        kw = dict(noisy_delay=True, noisy_delay_threshold=0.5, delay_mask=g["delay_mask"])
    obs = env.reset(customer_demand=np.array(g["demand_trace"]), **kw)
    if multi:
        assert list(obs.keys()) == names and obs[names[0]].shape == (env.obs_len,) and obs[names[0]].dtype == np.float64
        np.testing.assert_array_equal(np.stack([obs[n] for n in names]), ref["obs"][0])
    else:
        assert obs.shape == (m, env.obs_len)
        np.testing.assert_array_equal(obs, ref["obs"][0])
    assert env.period == 0
    for t in range(T):
        if multi:
            act = {names[i]: np.array([g["actions"][t, i]]) for i in range(m)}       # RLlib Box(1,) actions
        else:
            act = list(g["actions"][t])                                               # lists are accepted (np.squeeze)
        obs, rew, done, info = env.step(act)
        if multi:
            np.testing.assert_array_equal(np.stack([obs[n] for n in names]), ref["obs"][t + 1])
            for i, n in enumerate(names):
                assert isinstance(rew[n], np.float64) and rew[n] == ref["reward"][t, i]
                assert info[n]["period"] == t + 1
                assert info[n]["demand"] == ref["demand"][t, i] and info[n]["ship"] == ref["ship"][t, i]
                assert info[n]["acquisition"] == ref["acq"][t, i] and info[n]["actual order"] == ref["order"][t, i]
                assert info[n]["profit"] == ref["profit"][t, i]
            assert done == {"__all__": t == T - 1}
        else:
            np.testing.assert_array_equal(obs, ref["obs"][t + 1])
            assert rew == ref["reward"][t, 0]
            assert info["period"] == t
            np.testing.assert_array_equal(info["demand"], ref["demand"][t])
            np.testing.assert_array_equal(info["ship"], ref["ship"][t])
            np.testing.assert_array_equal(info["acquisition"], ref["acq"][t])
            np.testing.assert_array_equal(info["profit"], ref["profit"][t])
            assert done == (t == T - 1)
        assert env.period == t + 1
    # history arrays the scripts (and base_stock_policy) index
    np.testing.assert_array_equal(env.inv, ref["inv"])
    np.testing.assert_array_equal(env.backlog, ref["backlog"])
    np.testing.assert_array_equal(env.order_u, ref["order_u"])
    np.testing.assert_array_equal(env.order_r, ref["order"])
    with pytest.raises(IndexError):
        env.step(act)                                            # stepping past the end, like the reference's arrays


def test_attribute_surface_and_spaces():
    from marl_for_im_b200 import presets
    env = ENV_CLASSES["MAIM"](presets.serial4())
    assert env.num_agents == 4 and env.num_stages == 4 and env.num_nodes == 4 and env.num_periods == 30
    assert env.max_delay == 3 and env.prev_length == 1 and (env.a, env.b) == (-1, 1)
    np.testing.assert_array_equal(env.order_max, [30, 30, 30, 30])
    assert env.observation_space.shape == (7,) and env.action_space.shape == (1,)
    assert float(env.rescale(15, 0, 30, -1, 1)) == 0.0 and float(env.rev_scale(0.0, 0, 30, -1, 1)) == 15.0
    assert env.dist.name == "poisson" and env.dist_param == {"mu": 5}
    env = ENV_CLASSES["MAIM_div"](presets.div2(share_network=True))
    assert env.retailers == [3, 4, 5] and env.num_stages == 4 and env.observation_space.shape == (7,)
    np.testing.assert_array_equal(env.demand_max, [30, 60, 60, 30, 30, 30])
    np.testing.assert_array_equal(env.node_price, [2, 3, 4, 4, 5, 5])
    np.testing.assert_array_equal(env._demand_max_lib, env.demand_max)         # library and host derivations agree
    env = ENV_CLASSES["IM"](presets.serial4_dfo())
    assert env.observation_space.shape == (4, 3) and env.action_space.dtype == np.int32
    with pytest.raises(Exception, match="Not Implemented"):
        ENV_CLASSES["MAIM"](presets.serial4(time_dependency=False, prev_demand=False, prev_actions=True))
    with pytest.raises(ValueError):
        ENV_CLASSES["MAIM"](dict(presets.serial4(), delay=np.array([1, 0, 2, 1])))


def test_host_demand_draw_is_the_reference_stream():
    """reset() without a trace in drop-in mode uses scipy on the global numpy stream, seeded by the
    constructor like MAIM_env.py:50 — including quirk 4 (the serial classes pop mu on the first draw)."""
    from scipy.stats import poisson
    from marl_for_im_b200 import presets
    cfg = presets.serial4(mu=20)
    env = ENV_CLASSES["MAIM"](dict(cfg))                     # ctor: seed(52) then reset() draws with mu=20
    np.random.seed(52)
    want_first = poisson.rvs(size=30, mu=20)
    np.testing.assert_array_equal(env.customer_demand, want_first)
    state = np.random.get_state()
    env.reset()                                              # second draw: mu was popped → falls back to 5
    np.random.set_state(state)
    want_second = poisson.rvs(size=30, mu=5)
    np.testing.assert_array_equal(env.customer_demand, want_second)
    env_div = ENV_CLASSES["MAIM_div"](presets.div1(mu=20))   # divergent classes use .get: mu stays 20
    np.random.seed(52)
    np.testing.assert_array_equal(env_div.customer_demand, poisson.rvs(size=(2, 30), mu=20))
    state = np.random.get_state()
    env_div.reset()
    np.random.set_state(state)
    np.testing.assert_array_equal(env_div.customer_demand, poisson.rvs(size=(2, 30), mu=20))


def test_watchdog_sets_error_flag():
    """The reference raises 'Infinite Loop 4' when every child already holds its share and goods remain
    (reachable through a negative ledger); the batched env flags the env instead, the drop-in raises."""
    import torch
    from marl_for_im_b200 import presets
    cfg = presets.div1()
    env = ENV_CLASSES["MAIM_div"](dict(cfg, num_envs=8))
    env.reset(customer_demand=np.full((8, 2, 30), 5, dtype=np.int32))
    st = env.state_dict()
    st["backlog_to"][3] = torch.tensor([-40, -40], dtype=torch.int32, device="cuda:0")    # env 3: ledger far below zero
    st["inv"][3, 1] = 3                                                                     # node 1 can ship only part of the demand
    act = torch.zeros((8, 4), dtype=torch.float64, device="cuda:0")
    env.step(act)
    flags = env.error_flags.cpu().numpy()
    assert flags[3] == 4 and flags[[0, 1, 2, 4, 5, 6, 7]].sum() == 0


HOSTSTREAM_DIR = __import__("os").path.join(__import__("harness").GOLDEN_DIR, "hoststream")


@pytest.mark.parametrize("name", sorted(f[:-4] for f in __import__("os").listdir(HOSTSTREAM_DIR) if f.endswith(".npz")))
def test_dropin_consumes_the_host_random_stream_like_the_reference(name):
    """Episodes driven only by the constructor seed: the demand draw inside reset(), the noisy-demand
    mutations and the per-period noisy-delay uniforms must come off numpy's global stream in the
    reference's order (fixtures: tests/golden/make_golden_hoststream.py)."""
    import json
    z = np.load(__import__("os").path.join(HOSTSTREAM_DIR, name + ".npz"), allow_pickle=False)
    raw = json.loads(str(z["config"]))
    cfg = {}
    for k, v in raw.items():
        if isinstance(v, dict) and "__nd__" in v:
            cfg[k] = np.array(v["__nd__"], dtype=v["dtype"])
        elif isinstance(v, dict) and "__dict__" in v:
            cfg[k] = {int(a): list(b) for a, b in v["__dict__"].items()}
        elif isinstance(v, list):
            cfg[k] = tuple(v)
        else:
            cfg[k] = v
    kind = str(z["kind"])
    multi = kind.startswith("MAIM")
    thr = float(z["noisy_delay_threshold"])
    env = ENV_CLASSES[kind](cfg)
    m, T = env.num_nodes, env.num_periods
    names = agent_names(kind, m)
    for e in range(z["demands"].shape[0]):
        obs = env.reset(noisy_delay=True, noisy_delay_threshold=thr) if thr >= 0 else env.reset()
        np.testing.assert_array_equal(np.asarray(env.customer_demand), z["demands"][e], err_msg=f"{name} demand draw, episode {e}")
        got = np.stack([obs[n] for n in names]) if multi else obs
        np.testing.assert_array_equal(got, z["obs"][e, 0])
        for t in range(T):
            act = {names[i]: np.array([z["actions"][e, t, i]]) for i in range(m)} if multi else z["actions"][e, t]
            obs, rew, done, info = env.step(act)
            got = np.stack([obs[n] for n in names]) if multi else obs
            np.testing.assert_array_equal(got, z["obs"][e, t + 1], err_msg=f"{name} obs episode {e} t={t}")
            if multi:
                np.testing.assert_array_equal([rew[n] for n in names], z["reward"][e, t])
            else:
                assert rew == z["reward"][e, t, 0]
