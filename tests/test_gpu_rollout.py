"""Fused base-stock rollout kernel vs the oracle's step-by-step loop, the drop-in API mirrors of
base_restock_policy.py, and the Philox demand stream."""
import numpy as np
import pytest
import torch

from harness import run_oracle
from marl_for_im_b200 import presets
from oracle import im_oracle

pytestmark = pytest.mark.gpu


def oracle_rollout(kind, cfg, z, demand):
    env = im_oracle.OracleEnv(kind, dict(cfg))
    rewards = im_oracle.base_stock_rollout(env, z, demand)
    return env, rewards


@pytest.mark.parametrize("kind,preset", [("IM", "serial4_dfo"), ("MAIM", "serial8"), ("IM", "serial8"), ("MAIM", "serial2")])
def test_rollout_matches_oracle_serial(kind, preset):
    from marl_for_im_b200.envs import ENV_CLASSES
    cfg = presets.PRESETS[preset]()
    cfg.update(time_dependency=False, prev_demand=False, prev_actions=False, standardise_state=False, standardise_actions=False)
    m, T, N = cfg["num_stages"], 30, 257
    rng = np.random.default_rng(3)
    demand = rng.poisson(5, size=(N, T)).astype(np.int32)
    z = rng.integers(5, 41, size=(N, m)).astype(np.float64)
    z[::7] += 0.37                                              # Powell proposes non-integer levels
    env = ENV_CLASSES[kind](dict(cfg, num_envs=N))
    out = env.rollout_basestock(z, customer_demand=demand, step_rewards=True, write_state=True)
    ret, sr = out["returns"].cpu().numpy(), out["step_rewards"].cpu().numpy()
    st = {k: v.cpu().numpy() for k, v in env.state_dict().items()}
    assert env.period == T
    for n in range(0, N, 13):
        oenv, rewards = oracle_rollout(kind, cfg, z[n], demand[n])
        want_state = oenv.state_vector()
        for k in ("inv", "backlog", "order_u", "pipe"):
            np.testing.assert_array_equal(st[k][n], want_state[k], err_msg=k)
        if kind == "IM":
            np.testing.assert_array_equal(sr[:, n], np.array(rewards))
            acc = 0
            for r in rewards:
                acc += r
            assert ret[n] == acc
        else:
            np.testing.assert_array_equal(sr[:, n], np.stack(rewards))
            acc = np.zeros(m)
            for r in rewards:
                acc = acc + r
            np.testing.assert_array_equal(ret[n], acc)


def test_rollout_matches_oracle_div():
    from marl_for_im_b200.envs import MultiAgentInvManagementDiv, InvManagementDiv
    rng = np.random.default_rng(4)
    for cls, kind in ((MultiAgentInvManagementDiv, "MAIM_div"), (InvManagementDiv, "IM_div")):
        cfg = presets.div2(time_dependency=False, prev_demand=False, prev_actions=False)
        cfg["standardise_actions"] = False          # honoured by IM_div only; MAIM_div always rescales (quirk 9)
        N, T, m, R = 129, 30, 6, 3
        demand = rng.poisson(6, size=(N, R, T)).astype(np.int32)
        z = rng.integers(5, 41, size=(N, m)).astype(np.float64)
        env = cls(dict(cfg, num_envs=N))
        out = env.rollout_basestock(z, customer_demand=demand, step_rewards=True, write_state=True)
        sr = out["step_rewards"].cpu().numpy()
        st = {k: v.cpu().numpy() for k, v in env.state_dict().items()}
        for n in range(0, N, 11):
            oenv, rewards = oracle_rollout(kind, cfg, z[n], demand[n])
            want = oenv.state_vector()
            for k in ("inv", "backlog", "order_u", "pipe", "backlog_to"):
                np.testing.assert_array_equal(st[k][n], want[k], err_msg=f"{kind} {k}")
            got = sr[:, n] if kind == "IM_div" else sr[:, n, :]
            np.testing.assert_array_equal(got, np.array(rewards))


def test_dfo_func_and_dropin_basestock_loop():
    """The reference call pattern of inv_management.py:217-232 on the drop-in (N=1, numpy) env."""
    from scipy.stats import poisson
    from marl_for_im_b200.base_restock_policy import base_stock_policy, dfo_func
    from marl_for_im_b200.envs import InvManagement
    cfg = presets.serial4_dfo()
    env = InvManagement(dict(cfg))
    rng = np.random.default_rng(9)
    orc = im_oracle.OracleEnv("IM", cfg)
    for z in (np.array([25., 25., 25., 25.]), np.array([12.5, 31.25, 8.0, 40.0])):
        demand = rng.poisson(5, 30)
        want = im_oracle.dfo_value(orc, z, demand, poisson.pmf(demand, mu=5))
        assert dfo_func(z, env, demand) == want
        # step-by-step loop through reset()/step() with the policy reading env.inv[env.period, :]
        rewards = im_oracle.base_stock_rollout(orc, z, demand)
        env.reset(customer_demand=demand)
        done, t = False, 0
        while not done:
            s, r, done, info = env.step(base_stock_policy(z, env))
            assert r == rewards[t]
            assert info["period"] == t
            t += 1
        assert t == 30
        np.testing.assert_array_equal(env.inv[30], np.array(orc.inv, dtype=float))


def test_batched_basestock_policy_tensor_path():
    from marl_for_im_b200.base_restock_policy import base_stock_policy
    from marl_for_im_b200.envs import InvManagement
    cfg = presets.serial4_dfo()
    N, T = 64, 30
    rng = np.random.default_rng(2)
    demand = rng.poisson(5, size=(N, T)).astype(np.int32)
    z = np.array([25., 20., 30., 15.])
    env = InvManagement(dict(cfg, num_envs=N))
    env.reset(customer_demand=demand)
    total = torch.zeros(N, dtype=torch.float64, device="cuda:0")
    for t in range(T):
        _, r, done, _ = env.step(base_stock_policy(z, env))
        total += r
    fused = env.rollout_basestock(z, customer_demand=demand)["returns"]
    assert torch.equal(total, fused)


def philox_numpy(seed, env, tag, idx, t, episode):
    """numpy restatement of philox_draw (imx_device.cuh) for bit-exact stream checks."""
    M = 0xFFFFFFFF
    c = [env & M, (env >> 32) & M, ((tag << 28) | ((idx & 0xFFF) << 16) | ((t >> 1) & 0xFFFF)) & M, episode & M]
    k0, k1 = seed & M, ((seed >> 32) ^ (episode >> 32)) & M
    for _ in range(10):
        p0, p1 = 0xD2511F53 * c[0], 0xCD9E8D57 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k0) & M, p1 & M, ((p0 >> 32) ^ c[3] ^ k1) & M, p0 & M]
        k0, k1 = (k0 + 0x9E3779B9) & M, (k1 + 0xBB67AE85) & M
    return c


def test_philox_demand_stream_exact_and_shard_invariant():
    import ctypes as C
    from marl_for_im_b200.envs import MultiAgentInvManagement
    cfg = presets.serial4(mu=5)
    N = 1024
    env = MultiAgentInvManagement(dict(cfg, num_envs=N, seed=1234))
    env.reset()
    d = env.customer_demand_device().cpu().numpy()               # [T, 1, N]
    n_cdf = env._lib.imx_poisson_cdf(env._handle, None, 0)
    buf = (C.c_double * n_cdf)()
    env._lib.imx_poisson_cdf(env._handle, buf, n_cdf)
    cdf = np.array(buf[:])
    ep = env._episode
    for n in (0, 1, 17, 1023):
        for t in (0, 5, 28, 29):
            w = philox_numpy(1234, n, 0, 0, t, ep)
            hi, lo = (w[2], w[3]) if (t & 1) else (w[0], w[1])           # one Philox call serves periods 2k (x, y) and 2k+1 (z, w)
            u = float((((hi >> 5) << 26) | (lo >> 6))) / 9007199254740992.0
            assert d[t, 0, n] == int(np.searchsorted(cdf, u, side="right"))
    # sharding: envs [512, 1024) created as their own handle with env_offset draw the same trace
    e2 = MultiAgentInvManagement(dict(cfg, num_envs=512, env_offset=512, seed=1234))
    e2._episode = ep - 1
    e2.reset()
    np.testing.assert_array_equal(e2.customer_demand_device().cpu().numpy()[:, :, :], d[:, :, 512:])
    # distribution: chi-square against scipy's pmf
    from scipy.stats import poisson, chisquare
    big = MultiAgentInvManagement(dict(cfg, num_envs=200000, seed=7))
    big.reset()
    x = big.customer_demand_device().cpu().numpy().reshape(-1)
    kmax = 16
    obs = np.bincount(np.minimum(x, kmax), minlength=kmax + 1).astype(float)
    p = poisson.pmf(np.arange(kmax + 1), 5)
    p[kmax] = 1 - p[:kmax].sum()
    assert chisquare(obs, p * obs.sum()).pvalue > 1e-4
    assert abs(x.mean() - 5) < 0.01


def test_return_stats():
    from marl_for_im_b200.envs import MultiAgentInvManagement
    cfg = presets.serial4()
    N = 5000
    env = MultiAgentInvManagement(dict(cfg, num_envs=N))
    ret = torch.randn((N, 4), dtype=torch.float64, device="cuda:0")
    st = env.return_stats(ret).cpu().numpy()
    tot = ret.sum(dim=1)
    assert st[0] == N
    np.testing.assert_allclose(st[1], tot.sum().item(), rtol=1e-12)
    np.testing.assert_allclose(st[2], (tot * tot).sum().item(), rtol=1e-12)
    for i in range(4):
        np.testing.assert_allclose(st[3 + 2 * i], ret[:, i].sum().item(), rtol=1e-12)
        np.testing.assert_allclose(st[4 + 2 * i], (ret[:, i] ** 2).sum().item(), rtol=1e-12)


def test_episode_stats_accumulates_like_the_eval_loops():
    from marl_for_im_b200.envs import InvManagement, MultiAgentInvManagement
    N, T = 3001, 30
    env = MultiAgentInvManagement(dict(presets.serial4(), num_envs=N))
    sr = torch.randn((T, N, 4), dtype=torch.float64, device="cuda:0")
    st = env.episode_stats(sr)
    ret = torch.zeros((N, 4), dtype=torch.float64, device="cuda:0")
    for t in range(T):
        ret += sr[t]                                         # "reward += r" in period order
    want = env.return_stats(ret)
    assert torch.equal(st, want)
    st2 = env.episode_stats(sr * 2, stats=st.clone(), accumulate=True)
    np.testing.assert_allclose(st2.cpu().numpy()[0], 2 * N)
    np.testing.assert_allclose(st2.cpu().numpy()[1], 3 * want.cpu().numpy()[1], rtol=1e-12)
    np.testing.assert_allclose(st2.cpu().numpy()[2], 5 * want.cpu().numpy()[2], rtol=1e-12)
    env1 = InvManagement(dict(presets.serial4_dfo(), num_envs=N))
    sr1 = torch.randn((T, N), dtype=torch.float64, device="cuda:0")
    st1 = env1.episode_stats(sr1).cpu().numpy()
    tot = sr1.sum(dim=0)
    np.testing.assert_allclose(st1, [N, tot.sum().item(), (tot * tot).sum().item()], rtol=1e-12)


@pytest.mark.parametrize("kind,preset,N", [("MAIM", "serial2", 257), ("MAIM", "serial8", 4097), ("MAIM_div", "div2", 1000),
                                           ("MAIM_div", "div1", 33), ("IM", "serial4", 513), ("IM_div", "div2", 70000)])
def test_episode_statistics_every_slice_shape(kind, preset, N):
    """The two-stage statistics (one block per slice of envs, one warp per statistic pair) for agent counts 1 / 2 / 4 / 6 / 8 —
    i.e. 256, 256, 256, 160 and 128 envs per slice — and env counts that end inside a slice: returns = step rewards added in
    period order (bit for bit), statistics against float64 numpy within summation-order tolerance, and the fused form equal to the
    plain form on its own returns."""
    from marl_for_im_b200.envs import ENV_CLASSES
    env = ENV_CLASSES[kind](dict(presets.PRESETS[preset](), num_envs=N))
    m, T = env.num_nodes, env.num_periods
    cols = m if env.MULTI else 1
    g = torch.Generator(device="cuda:0")
    g.manual_seed(N)
    sr = torch.randn((T, N, cols) if env.MULTI else (T, N), dtype=torch.float64, device="cuda:0", generator=g) * 50.0
    ret = torch.empty((N, cols) if env.MULTI else (N,), dtype=torch.float64, device="cuda:0")
    st = env.episode_stats(sr, returns=ret)
    want_ret = torch.zeros_like(ret)
    for t in range(T):
        want_ret += sr[t]
    assert torch.equal(ret, want_ret)
    assert torch.equal(st, env.return_stats(ret))
    r = want_ret.reshape(N, cols).cpu().numpy()
    tot = r.sum(axis=1)
    want = [float(N), tot.sum(), (tot * tot).sum()]
    if env.MULTI:
        for c in range(cols):
            want += [r[:, c].sum(), (r[:, c] ** 2).sum()]
    np.testing.assert_allclose(st.cpu().numpy(), want, rtol=1e-11, atol=1e-6)


def test_dfo_func_batch_matches_single_env_objective():
    from scipy.stats import poisson
    from marl_for_im_b200.base_restock_policy import dfo_func_batch, population_search_inventory_policy
    from marl_for_im_b200.envs import InvManagement
    cfg = presets.serial4_dfo()
    rng = np.random.default_rng(21)
    K, D = 7, 9
    policies = rng.integers(10, 40, size=(K, 4)).astype(float) + rng.choice([0.0, 0.25], size=(K, 4))
    demands = rng.poisson(5, size=(D, 30))
    env = InvManagement(dict(cfg, num_envs=K * D))
    got = dfo_func_batch(policies, env, demands).cpu().numpy()
    orc = im_oracle.OracleEnv("IM", cfg)
    for k in range(K):
        for d in range(D):
            assert got[k, d] == im_oracle.dfo_value(orc, policies[k], demands[d], poisson.pmf(demands[d], mu=5))
    pol, score = population_search_inventory_policy(InvManagement, cfg, demands, np.ones(4) * 25, sweeps=2, radius=3)
    base = dfo_func_batch(np.ones((1, 4)) * 25, InvManagement(dict(cfg, num_envs=D)), demands).mean().item()
    assert pol.shape == (4,) and score <= base + 1e-12


def test_device_noisy_demand_generator():
    """noisy_demand (MAIM_div_env.py:287-295) on the device: each generated demand is doubled with probability thr and
    then zeroed with probability thr, from its own Philox tag; replayed traces are never touched; the fused rollout's
    in-kernel generator draws the same trace."""
    from marl_for_im_b200.envs import InvManagementDiv, MultiAgentInvManagementDiv
    thr, N, seed = 0.2, 8192, 99
    cfg = presets.div2()
    clean = MultiAgentInvManagementDiv(dict(cfg, num_envs=N, seed=seed))
    clean.reset()
    d0 = clean.customer_demand_device().cpu().numpy()                     # [T, R, N]
    noisy = MultiAgentInvManagementDiv(dict(cfg, num_envs=N, seed=seed, noisy_demand=True, noisy_demand_threshold=thr))
    noisy.reset()
    d1 = noisy.customer_demand_device().cpu().numpy()
    ep = noisy._episode
    assert ep == clean._episode
    T, R, _ = d0.shape
    u = lambda hi, lo: float((((hi >> 5) << 26) | (lo >> 6))) / 9007199254740992.0   # noqa: E731
    for n in (0, 3, 1000, N - 1):
        for r in range(R):
            for t in range(T):
                w = philox_numpy(seed, n, 2, r, 2 * t, ep)                # tag 2, counter = the period
                want = int(d0[t, r, n])
                if u(w[0], w[1]) <= thr:
                    want *= 2
                if u(w[2], w[3]) <= thr:
                    want = 0
                assert d1[t, r, n] == want
    zero_rate = (d1 == 0).mean()
    p0 = np.exp(-5.0)
    assert abs(zero_rate - (thr + (1 - thr) * p0)) < 0.005
    doubled = (d1 == 2 * d0) & (d0 > 0)
    assert abs(doubled.mean() / (1 - p0) - thr * (1 - thr)) < 0.01
    # replayed demand is left alone
    trace = np.random.default_rng(1).poisson(5, size=(N, R, T)).astype(np.int32)
    noisy.reset(customer_demand=trace)
    np.testing.assert_array_equal(noisy.customer_demand_device().cpu().numpy(), trace.transpose(2, 1, 0))
    # the fused rollout draws the same noisy trace in-kernel as reset() stored for the step path
    icfg = presets.div2()
    env = InvManagementDiv(dict(icfg, num_envs=N, seed=seed, noisy_demand=True, noisy_demand_threshold=thr))
    env.reset()
    stored = env.customer_demand_device().permute(2, 1, 0).contiguous()  # [N, R, T]
    z = np.array([40., 35., 20., 18., 12., 15.])
    ep = env._episode
    a = env.rollout_basestock(z, customer_demand=stored)["returns"]
    env._episode = ep - 1
    b = env.rollout_basestock(z)["returns"]
    assert torch.equal(a, b)
