"""Centralised-critic observer: oracle restatement pinned to the reference function (CPU), and the
CUDA kernel against the oracle (GPU)."""
import numpy as np
import pytest

from harness import reference_available
from oracle import im_oracle


@pytest.mark.skipif(not reference_available(), reason="reference tree not present")
def test_oracle_cc_matches_reference_function():
    from oracle.ref_import import load_reference_cc_observer
    ref = load_reference_cc_observer()
    rng = np.random.default_rng(0)
    for m, O in ((2, 8), (4, 7), (8, 8), (3, 5)):
        agent_obs = {f"stage_{i}": rng.uniform(-1, 1, O) for i in range(m)}
        want = ref(agent_obs)
        got = im_oracle.central_critic_flat(np.stack([agent_obs[f"stage_{i}"] for i in range(m)]))
        for i in range(m):
            w = want[f"stage_{i}"]
            flat = np.concatenate([w["opponent_action"], w["opponent_obs"], w["own_obs"]])     # sorted-key flattening
            np.testing.assert_array_equal(got[i], flat)


@pytest.mark.gpu
def test_cuda_cc_observer_matches_oracle():
    import torch
    from marl_for_im_b200 import presets
    from marl_for_im_b200.cc import cc_observe, central_critic_observer
    from marl_for_im_b200.envs import MultiAgentInvManagement
    rng = np.random.default_rng(1)
    for cfg in (presets.serial2(), presets.serial4(), presets.serial8(prev_actions=True)):
        N = 777
        env = MultiAgentInvManagement(dict(cfg, num_envs=N))
        m, O = env.num_nodes, env.obs_len
        demand = rng.poisson(5, size=(N, 30)).astype(np.int32)
        obs = env.reset(customer_demand=demand)
        act = torch.as_tensor(rng.uniform(-1.3, 1.3, size=(N, m)), device="cuda:0")
        obs, _, _, _ = env.step(act)
        full = torch.stack([obs[n] for n in env.agent_names], dim=1)
        got0 = cc_observe(env, obs).cpu().numpy()
        got1 = cc_observe(env, full, actions=act).cpu().numpy()
        got32 = cc_observe(env, full, actions=act, dtype=torch.float32).cpu().numpy()
        full_h, act_h = full.cpu().numpy(), act.cpu().numpy()
        for n in (0, 1, 400, N - 1):
            np.testing.assert_array_equal(got0[n], im_oracle.central_critic_flat(full_h[n]))
            np.testing.assert_array_equal(got1[n], im_oracle.central_critic_flat(full_h[n], act_h[n]))
            np.testing.assert_array_equal(got32[n], im_oracle.central_critic_flat(full_h[n], act_h[n]).astype(np.float32))
        d = central_critic_observer(obs, env=env)
        assert set(d["stage_0"]) == {"own_obs", "opponent_obs", "opponent_action"}
        assert d["stage_1"]["own_obs"].shape == (N, O) and d["stage_1"]["opponent_obs"].shape == (N, (m - 1) * O)
        assert torch.equal(d["stage_1"]["own_obs"], obs["stage_1"])
    # an env that writes float32 observations: the observer reads them as float32 (same values as casting afterwards)
    cfg = presets.serial2()
    e64 = MultiAgentInvManagement(dict(cfg, num_envs=512))
    e32 = MultiAgentInvManagement(dict(cfg, num_envs=512, obs_dtype="float32"))
    demand = rng.poisson(5, size=(512, 30)).astype(np.int32)
    act = torch.as_tensor(rng.uniform(-1, 1, size=(512, 2)), device="cuda:0")
    for e in (e64, e32):
        e.reset(customer_demand=demand)
        e.step(act)
    want = cc_observe(e64, e64.last_obs, actions=act, dtype=torch.float32)
    got = cc_observe(e32, e32.last_obs, actions=act, dtype=torch.float32)
    assert e32.last_obs.dtype == torch.float32 and torch.equal(got, want)
    # drop-in (N = 1, numpy dicts) — the reference's call pattern
    env = MultiAgentInvManagement(presets.serial2())
    o = env.reset(customer_demand=np.full(30, 5))
    d = central_critic_observer(o, env=env)
    np.testing.assert_array_equal(d["stage_0"]["own_obs"], o["stage_0"])
    np.testing.assert_array_equal(d["stage_0"]["opponent_obs"], o["stage_1"])
    np.testing.assert_array_equal(d["stage_0"]["opponent_action"], np.zeros(1))


@pytest.mark.gpu
@pytest.mark.parametrize("preset,N,obs_dtype,pipe", [("serial2", 65536, "float32", "0"), ("serial2", 65536, "float32", "1"), ("serial4", 4096, "float64", "0"),
                                                     ("serial8", 8192, "float64", "1"), ("serial4", 4100, "float64", "0"), ("serial2", 96, "float32", "0")])
def test_fused_cc_rows_from_the_step_kernel(preset, N, obs_dtype, pipe, monkeypatch):
    """imx_step_cc: the critic rows the step kernel emits == oracle.central_critic_flat of the oracle's observations, and
    == the separate gather kernel; fused variants (one tile per CTA, pipelined) and the unfused fallbacks (tail, N < 1024)."""
    import torch
    from marl_for_im_b200 import presets
    from marl_for_im_b200.cc import cc_observe
    from marl_for_im_b200.envs import MultiAgentInvManagement
    from oracle import c_oracle
    monkeypatch.setenv("IMX_PIPE", pipe)
    if pipe == "1":
        monkeypatch.setenv("IMX_PIPE_CTAS", "1")
    cfg = presets.PRESETS[preset]()
    rng = np.random.default_rng(3)
    env = MultiAgentInvManagement(dict(cfg, num_envs=N, obs_dtype=obs_dtype))
    m, O, T = env.num_nodes, env.obs_len, 6
    demand = rng.poisson(5, size=(N, 1, 30)).astype(np.int32)
    actions = rng.uniform(-1.3, 1.3, size=(T, N, m))
    env.reset(customer_demand=demand)
    a_dev = torch.as_tensor(actions, device="cuda:0")
    np_dt = np.float32 if obs_dtype == "float32" else np.float64
    want = c_oracle.COracle("MAIM", cfg).run(demand, actions, periods=T, all_obs=True)
    for t in range(T):
        fill = bool(t % 2 == 0)
        obs, cc, rew, done = env.step_cc(a_dev[t], fill_actions=fill, want_obs=(t != 3))
        cc_h = cc.cpu().numpy()
        assert cc_h.dtype == np_dt
        if obs is not None:
            np.testing.assert_array_equal(obs.cpu().numpy(), want["obs_all"][t + 1].astype(np_dt))
            sep = cc_observe(env, obs, actions=a_dev[t] if fill else None, dtype=env.obs_dtype).cpu().numpy()
            np.testing.assert_array_equal(cc_h, sep)
        np.testing.assert_array_equal(rew.cpu().numpy(), want["reward"][t])
        for n in (0, 1, N // 2, N - 1):
            flat = im_oracle.central_critic_flat(want["obs_all"][t + 1, n], actions[t, n] if fill else None)
            np.testing.assert_array_equal(cc_h[n], flat.astype(np_dt), err_msg=f"t={t} n={n}")
    if N >= 1024 and N % 64 == 0:        # fused: a specialised kernel served it (the pipelined one when its ring of critic-row tiles fits in shared memory)
        assert env._lib.imx_kernel_variant(env._handle) in ((2, 3) if pipe == "1" else (2,))
