"""CUDA path vs the oracle / the reference-generated golden fixtures — serial chains.
Every call goes through the C ABI (ctypes → libimx_b200.so); bit-exact on all fields."""
import itertools

import numpy as np
import pytest
import torch

from harness import golden_names, load_golden, random_case, run_oracle, make_delay_mask
from marl_for_im_b200 import presets

pytestmark = pytest.mark.gpu


from cuda_harness import assert_same, run_cuda  # noqa: E402


SERIAL_GOLDEN = [n for n in golden_names() if not ("div" in n)]


@pytest.mark.parametrize("name", SERIAL_GOLDEN)
def test_cuda_matches_golden_serial(name):
    g = load_golden(name)
    got = run_cuda(g["kind"], g["config"], g["demand_trace"], g["actions"], g["delay_mask"], n_copies=72)
    assert_same(g["ref"], got, name)


MODES = list(itertools.product([False, True], repeat=3))


@pytest.mark.parametrize("kind,preset", [("MAIM", "serial4"), ("IM", "serial4"), ("MAIM", "serial8"), ("IM", "serial8"),
                                         ("MAIM", "serial2")])
def test_cuda_matches_oracle_all_modes(kind, preset):
    rng = np.random.default_rng(99)
    for td, pd, pa in MODES:
        if kind == "MAIM" and (not td) and pa and (not pd):
            continue
        for P, mu, indep in [(1, 5, False), (3, 15, True)]:
            cfg = presets.PRESETS[preset](time_dependency=td, prev_demand=pd, prev_actions=pa, prev_length=P, independent=indep)
            m = cfg["num_stages"]
            cfg["inv_max"] = np.array([30, 25, 40, 35, 30, 45, 20, 30][:m], dtype=float)
            cfg["inv_target"] = np.array([0, 3, 5.5, 1, 0, 2, 4, 0][:m], dtype=float)
            demand, actions = random_case(kind, cfg, rng, mu=mu)
            assert_same(run_oracle(kind, cfg, demand, actions), run_cuda(kind, cfg, demand, actions, n_copies=72),
                        f"{kind}/{preset}/{(td, pd, pa)}/P{P}")


@pytest.mark.parametrize("kind", ["MAIM", "IM"])
def test_cuda_non_standardised_and_noisy(kind):
    rng = np.random.default_rng(5)
    for td, pd, pa in MODES:
        if kind == "MAIM" and (not td) and pa and (not pd):
            continue
        cfg = presets.serial4(time_dependency=td, prev_demand=pd, prev_actions=pa, prev_length=2,
                              standardise_state=False, standardise_actions=False)
        demand, actions = random_case(kind, cfg, rng, mu=12)
        assert_same(run_oracle(kind, cfg, demand, actions), run_cuda(kind, cfg, demand, actions, n_copies=72), "raw")
    cfg = presets.serial8()
    for _ in range(3):
        demand, actions = random_case(kind, cfg, rng, mu=6, action_mode="near_eq")
        mask = make_delay_mask(kind, cfg["delay"], 30, 0.3, rng)
        assert_same(run_oracle(kind, cfg, demand, actions, mask), run_cuda(kind, cfg, demand, actions, mask, n_copies=72), "noisy")


def test_cuda_wide_chain_and_tail_sizes():
    """m not a power of two, m up to 32 (one warp per env), N not a multiple of the warp tile."""
    rng = np.random.default_rng(17)
    for m in (3, 5, 11, 16, 23, 32):
        cfg = {
            "num_stages": m, "num_periods": 12, "init_inv": np.ones(m) * 8, "inv_target": np.ones(m) * 2,
            "inv_max": rng.integers(15, 40, m).astype(float), "price": np.arange(m + 1, 0, -1).astype(float),
            "stock_cost": rng.uniform(0.1, 0.5, m), "backlog_cost": rng.uniform(0.3, 0.9, m),
            "delay": rng.integers(1, 6, m), "time_dependency": True, "prev_demand": True, "prev_actions": True,
            "prev_length": 2, "independent": False,
        }
        for kind in ("MAIM", "IM"):
            demand = rng.poisson(6, 12)
            actions = rng.uniform(-1.1, 1.1, (12, m))
            assert_same(run_oracle(kind, cfg, demand, actions), run_cuda(kind, cfg, demand, actions, n_copies=72), f"m={m}")


def test_cuda_batch_of_distinct_envs():
    """4096 envs with distinct traces, compared env by env on a sample and through shard invariance."""
    from marl_for_im_b200.envs import MultiAgentInvManagement
    cfg = presets.serial4()
    N, T, m = 4096 + 8, 30, 4
    rng = np.random.default_rng(420)
    demand = rng.poisson(5, size=(N, T)).astype(np.int32)
    actions = np.random.default_rng(0).uniform(-1, 1, size=(T, N, m))
    env = MultiAgentInvManagement(dict(cfg, num_envs=N))
    o = env.reset(customer_demand=demand)
    obs, rew = [], []
    a_dev = torch.as_tensor(actions, device="cuda:0")
    for t in range(T):
        o, r, done, _ = env.step(a_dev[t])
        obs.append(torch.stack([o[n] for n in env.agent_names], dim=1).cpu().numpy())
        rew.append(torch.stack([r[n] for n in env.agent_names], dim=1).cpu().numpy())
    obs, rew = np.stack(obs), np.stack(rew)
    for n in list(range(0, N, 97)) + [N - 1]:
        want = run_oracle("MAIM", cfg, demand[n], actions[:, n])
        np.testing.assert_array_equal(want["obs"][1:], obs[:, n])
        np.testing.assert_array_equal(want["reward"], rew[:, n])
    # shard invariance: the same envs split over two handles give the same bytes
    half = N // 2
    for lo, hi in ((0, half), (half, N)):
        e2 = MultiAgentInvManagement(dict(cfg, num_envs=hi - lo, env_offset=lo))
        e2.reset(customer_demand=demand[lo:hi])
        for t in range(T):
            o2, r2, _, _ = e2.step(a_dev[t, lo:hi])
        np.testing.assert_array_equal(torch.stack([o2[n] for n in e2.agent_names], dim=1).cpu().numpy(), obs[-1, lo:hi])


def test_both_step_paths_agree(monkeypatch):
    """The TMA-staged kernel and the direct kernel are two implementations of the same step: force
    each one (IMX_STEP_PATH is read when a handle is created) and compare bytes, including an N that
    is not a multiple of 4 (TMA illegal → the auto path must fall back to the direct kernel)."""
    rng = np.random.default_rng(77)
    cfg = presets.serial8(prev_actions=True, prev_length=2)
    demand, actions = random_case("MAIM", cfg, rng, mu=7)
    want = run_oracle("MAIM", cfg, demand, actions)
    for path, n in (("direct", 96), ("tma", 96), ("auto", 96), ("auto", 97), ("tma", 32)):
        if path == "auto":
            monkeypatch.delenv("IMX_STEP_PATH", raising=False)
        else:
            monkeypatch.setenv("IMX_STEP_PATH", path)
        assert_same(want, run_cuda("MAIM", cfg, demand, actions, n_copies=n), f"path={path} N={n}")


def test_runtime_specialised_kernels_match(monkeypatch):
    """IMX_JIT=1 forces the NVRTC-specialised build of the same kernel sources; it must be the one
    that runs (kernel variant 2) and give the same bytes as the oracle."""
    from marl_for_im_b200.envs import ENV_CLASSES
    from harness import copy_config
    monkeypatch.setenv("IMX_JIT", "1")
    rng = np.random.default_rng(2024)
    cases = [("MAIM", presets.serial4()), ("MAIM", presets.serial8(prev_actions=True, prev_length=3, independent=True)),
             ("IM", presets.serial4(time_dependency=False, prev_actions=True)), ("MAIM", presets.serial2()),
             ("IM", presets.serial4_dfo()), ("MAIM_div", presets.div1()), ("MAIM_div", presets.div2(share_network=True, prev_actions=True)),
             ("IM_div", presets.div2(prev_length=2))]
    for kind, cfg in cases:
        amode = "near_eq" if kind.endswith("div") else "uniform"
        demand, actions = random_case(kind, cfg, rng, mu=6, action_mode=amode)
        want = run_oracle(kind, cfg, demand, actions)
        # through the public env API, without info buffers (the specialised kernel serves the plain step)
        c = copy_config(cfg)
        c.update(num_envs=128, return_info=False)
        env = ENV_CLASSES[kind](c)
        d = np.broadcast_to(np.asarray(demand)[None], (128,) + np.asarray(demand).shape)
        o = env.reset(customer_demand=d)
        multi = kind.startswith("MAIM")
        for t in range(env.num_periods):
            a = torch.as_tensor(np.broadcast_to(actions[t][None], (128, env.num_nodes)).copy(), device="cuda:0")
            o, r, done, _ = env.step(a)
            assert env._lib.imx_kernel_variant(env._handle) in (2, 3), env._lib.imx_jit_log()
            got_o = (torch.stack([o[n] for n in env.agent_names], dim=1) if multi else o).cpu().numpy()
            got_r = (torch.stack([r[n] for n in env.agent_names], dim=1) if multi else r[:, None]).cpu().numpy()
            for n in (0, 63, 64, 127):
                np.testing.assert_array_equal(got_o[n], want["obs"][t + 1], err_msg=f"{kind} obs t={t}")
                np.testing.assert_array_equal(got_r[n, :got_r.shape[1]], want["reward"][t, :got_r.shape[1]], err_msg=f"{kind} reward t={t}")
        st = env.state_dict()
        np.testing.assert_array_equal(st["inv"][5].cpu().numpy(), want["inv"][-1])
        np.testing.assert_array_equal(st["backlog"][77].cpu().numpy(), want["backlog"][-1])


def test_host_buffer_path_matches_device_path(monkeypatch):
    """imx_reset_host / imx_step_host on pinned buffers (zero-copy: the kernel addresses host memory),
    on pinned buffers with staged copies, and on pageable numpy memory — same bytes as the device path."""
    import ctypes as C
    from marl_for_im_b200 import _lib
    from marl_for_im_b200.envs import MultiAgentInvManagement
    cfg = presets.serial4()
    N, T, m = 4096 + 64, 30, 4
    rng = np.random.default_rng(6)
    demand = rng.poisson(5, size=(N, 1, T)).astype(np.int32)
    actions = rng.uniform(-1, 1, size=(T, N, m))
    ref_env = MultiAgentInvManagement(dict(cfg, num_envs=N))
    ref_env.reset(customer_demand=demand)
    want_obs, want_rew = [], []
    a_dev = torch.as_tensor(actions, device="cuda:0")
    for t in range(T):
        o, r, _, _ = ref_env.step(a_dev[t])
        want_obs.append(torch.stack([o[n] for n in ref_env.agent_names], dim=1).cpu().numpy())
        want_rew.append(torch.stack([r[n] for n in ref_env.agent_names], dim=1).cpu().numpy())
    for mode in ("zero_copy", "staged_pinned", "pageable"):
        monkeypatch.setenv("IMX_HOST_ZERO_COPY", "0" if mode == "staged_pinned" else "1")
        env = MultiAgentInvManagement(dict(cfg, num_envs=N))
        O = env.obs_len
        if mode == "pageable":
            dem_h, act_h = demand.copy(), actions.copy()
            obs_h, rew_h = np.empty((N, m, O)), np.empty((N, m))
            ptr = lambda a: C.c_void_p(a.ctypes.data)          # noqa: E731
            view = lambda a: a                                 # noqa: E731
        else:
            dem_h, act_h = torch.as_tensor(demand).pin_memory(), torch.as_tensor(actions).pin_memory()
            obs_h, rew_h = torch.empty((N, m, O), dtype=torch.float64).pin_memory(), torch.empty((N, m), dtype=torch.float64).pin_memory()
            ptr = lambda a: C.c_void_p(a.data_ptr())           # noqa: E731
            view = lambda a: a.numpy()                         # noqa: E731
        _lib.check(env._lib.imx_reset_host(env._handle, ptr(dem_h), None, 0, 3, ptr(obs_h)))
        for t in range(T):
            _lib.check(env._lib.imx_step_host(env._handle, ptr(act_h[t]), ptr(obs_h), ptr(rew_h)))
            np.testing.assert_array_equal(view(obs_h), want_obs[t], err_msg=f"{mode} obs t={t}")
            np.testing.assert_array_equal(view(rew_h), want_rew[t], err_msg=f"{mode} reward t={t}")


@pytest.mark.parametrize("kind,preset,n,obs_dtype", [("MAIM", "serial4", 65536, "float64"), ("MAIM", "serial4", 16384 + 4096, "float32"),
                                                      ("MAIM_div", "div2", 32768, "float64"), ("IM", "serial8", 24576, "float64")])
def test_host_buffer_path_large_batches(kind, preset, n, obs_dtype, monkeypatch):
    """Large batches on pinned buffers (zero-copy: the TMA kernel addresses host memory) and with staged copies must
    equal the device path, including the period counter and the past-the-end error."""
    import ctypes as C
    from marl_for_im_b200 import _lib
    from marl_for_im_b200.envs import ENV_CLASSES
    from harness import copy_config
    cfg = presets.PRESETS[preset]()
    ref_env = ENV_CLASSES[kind](dict(copy_config(cfg), num_envs=n, obs_dtype=obs_dtype))
    m, T, O, R = ref_env.num_nodes, ref_env.num_periods, ref_env.obs_len, len(ref_env._retailers)
    rng = np.random.default_rng(n)
    demand = rng.poisson(5, size=(n, R, T)).astype(np.int32)
    actions = np.clip(rng.normal(-0.4, 0.6, size=(T, n, m)), -1.1, 1.1)
    ref_env.reset(customer_demand=demand)
    want_obs, want_rew = [ref_env.last_obs.cpu().numpy()], []
    a_dev = torch.as_tensor(actions, device="cuda:0")
    for t in range(T):
        ref_env.step(a_dev[t])
        want_obs.append(ref_env.last_obs.cpu().numpy())
        want_rew.append(ref_env.last_reward.cpu().numpy())
    want_state = {k: v.cpu().numpy() for k, v in ref_env.state_dict().items()}
    odt = torch.float32 if obs_dtype == "float32" else torch.float64
    for pipeline in ("1", "0"):
        monkeypatch.setenv("IMX_HOST_ZERO_COPY", pipeline)
        env = ENV_CLASSES[kind](dict(copy_config(cfg), num_envs=n, obs_dtype=obs_dtype))
        dem_h, act_h = torch.as_tensor(demand).pin_memory(), torch.as_tensor(actions).pin_memory()
        obs_h = torch.empty((n, m, O), dtype=odt).pin_memory()
        rew_h = torch.empty((n, m) if env.MULTI else (n,), dtype=torch.float64).pin_memory()
        p = lambda a: C.c_void_p(a.data_ptr())   # noqa: E731
        _lib.check(env._lib.imx_reset_host(env._handle, p(dem_h), None, 0, 3, p(obs_h)))
        np.testing.assert_array_equal(obs_h.numpy(), want_obs[0])
        for t in range(T):
            obs_h.zero_()
            rew_h.zero_()
            _lib.check(env._lib.imx_step_host(env._handle, p(act_h[t]), p(obs_h), p(rew_h)))
            np.testing.assert_array_equal(obs_h.numpy(), want_obs[t + 1], err_msg=f"zero_copy={pipeline} obs t={t}")
            np.testing.assert_array_equal(rew_h.numpy(), want_rew[t], err_msg=f"zero_copy={pipeline} reward t={t}")
        assert env.period == T
        for k, v in env.state_dict().items():
            np.testing.assert_array_equal(v.cpu().numpy(), want_state[k], err_msg=k)
        with pytest.raises(IndexError):
            _lib.check(env._lib.imx_step_host(env._handle, p(act_h[0]), p(obs_h), p(rew_h)))


def test_long_episode_and_demand_layout():
    """150-period episodes on distinct envs: the stored trace is the
    caller's [N][R][T] tensor transposed to [T][R][N], and the dynamics match the oracle to the last period."""
    from marl_for_im_b200.envs import MultiAgentInvManagement, MultiAgentInvManagementDiv
    rng = np.random.default_rng(150)
    for cls, kind, cfg, R in ((MultiAgentInvManagement, "MAIM", presets.serial4(), 1), (MultiAgentInvManagementDiv, "MAIM_div", presets.div2(), 3)):
        cfg = dict(cfg, num_periods=150)
        N, T = 333, 150
        m = cfg.get("num_nodes", cfg.get("num_stages"))
        demand = rng.poisson(5, size=(N, R, T)).astype(np.int32)
        actions = np.clip(rng.normal(-0.5, 0.5, size=(T, N, m)), -1, 1)
        env = cls(dict(cfg, num_envs=N))
        env.reset(customer_demand=demand if R > 1 else demand[:, 0])
        np.testing.assert_array_equal(env.customer_demand_device().cpu().numpy(), demand.transpose(2, 1, 0))
        a_dev = torch.as_tensor(actions, device="cuda:0")
        for t in range(T):
            env.step(a_dev[t])
        obs = env.last_obs.cpu().numpy()
        for n in (0, 127, 128, 332):
            want = run_oracle(kind, cfg, demand[n] if R > 1 else demand[n, 0], actions[:, n])
            np.testing.assert_array_equal(obs[n], want["obs"][-1])


def test_handles_sharing_a_kernel_with_different_tile_sizes():
    """Two live handles whose configs map to the SAME kernel instantiation but need different amounts of shared memory
    (observation length / element size): the function's dynamic-shared-memory limit must only ever be raised."""
    from marl_for_im_b200.envs import MultiAgentInvManagement
    rng = np.random.default_rng(12)
    N = 512
    demand = rng.poisson(5, size=(N, 30)).astype(np.int32)
    act = torch.as_tensor(rng.uniform(-1, 1, size=(30, N, 2)), device="cuda:0")
    big = MultiAgentInvManagement(dict(presets.serial2(), num_envs=N))                                   # O = 8 float64
    small = MultiAgentInvManagement(dict(presets.serial2(prev_actions=False), num_envs=N, obs_dtype="float32"))   # created later, smaller tile
    ref = MultiAgentInvManagement(dict(presets.serial2(), num_envs=N + 1))                               # direct kernel (N % 4 != 0)
    d1 = np.concatenate([demand, demand[:1]])
    for e, d in ((big, demand), (small, demand), (ref, d1)):
        e.reset(customer_demand=d)
    for t in range(30):
        big.step(act[t])
        small.step(act[t])
        ref.step(torch.cat([act[t], act[t][:1]]))
        assert torch.equal(big.last_obs, ref.last_obs[:N]) and torch.equal(big.last_reward, ref.last_reward[:N])


@pytest.mark.parametrize("kind", ["MAIM", "IM"])
def test_cuda_random_serial_chains(kind, monkeypatch):
    """Random chains (2-12 stages, lead times up to 5, histories up to 4, rescale intervals other than [-1, 1], raw and
    standardised actions) against the oracle — pinned to the reference on the same generator by
    tests/test_oracle_vs_reference.py — through the ahead-of-time and the runtime-specialised kernels."""
    from harness import random_serial_config
    rng = np.random.default_rng(5050 if kind == "MAIM" else 5051)
    done = 0
    while done < 16:
        cfg = random_serial_config(rng, int(rng.integers(2, 13)))
        if kind == "MAIM" and (not cfg["time_dependency"]) and cfg["prev_actions"] and (not cfg["prev_demand"]):
            continue
        T, m = cfg["num_periods"], cfg["num_stages"]
        demand = rng.poisson(rng.uniform(3, 12), T)
        if cfg["standardise_actions"]:
            span = cfg["b"] - cfg["a"]
            actions = rng.uniform(cfg["a"] - 0.1 * span, cfg["b"] + 0.1 * span, size=(T, m))
        else:
            actions = rng.uniform(-3, 50, size=(T, m))
        want = run_oracle(kind, cfg, demand, actions)
        monkeypatch.setenv("IMX_JIT", "0")
        assert_same(want, run_cuda(kind, cfg, demand, actions, n_copies=72), f"{kind} {cfg}")
        # the runtime-specialised kernels serve steps without diagnostics: observations and rewards through them
        monkeypatch.setenv("IMX_JIT", "1")
        from harness import copy_config
        from marl_for_im_b200.envs import ENV_CLASSES
        env = ENV_CLASSES[kind](dict(copy_config(cfg), num_envs=64, return_info=False))
        env.reset(customer_demand=np.broadcast_to(demand[None], (64, T)))
        for t in range(T):
            env.step(torch.as_tensor(np.broadcast_to(actions[t][None], (64, m)).copy(), device="cuda:0"))
            assert env._lib.imx_kernel_variant(env._handle) in (2, 3), env._lib.imx_jit_log()
            np.testing.assert_array_equal(env.last_obs[37].cpu().numpy(), want["obs"][t + 1], err_msg=f"jit obs t={t} {cfg}")
            r = env.last_reward[37].cpu().numpy()
            np.testing.assert_array_equal(r if env.MULTI else np.array([r]), want["reward"][t][:m if env.MULTI else 1], err_msg=f"jit reward t={t}")
        done += 1
