"""The persistent pipelined step kernel (csrc/imx_step_pipe.cuh, kernel variant 3) against the plain-C oracle: every env of
batches whose tile count does and does not divide the grid, ring depths 2..4, 1..4 resident CTAs per SM (IMX_PIPE_CTAS = 1
with a few thousand envs makes every CTA walk many tiles), all four kinds, float32 observations, obs = NULL, and the tail
that falls to the direct kernel."""
import numpy as np
import pytest
import torch

from marl_for_im_b200 import presets
from marl_for_im_b200.envs import ENV_CLASSES
from oracle import c_oracle

pytestmark = pytest.mark.gpu


def _episode(kind, cfg, N, rng, periods=None, obs_dtype=None):
    extra = {} if obs_dtype is None else {"obs_dtype": obs_dtype}
    env = ENV_CLASSES[kind](dict(cfg, num_envs=N, **extra))
    m, T, R = env.num_nodes, env.num_periods, len(env._retailers)
    periods = periods or T
    demand = rng.poisson(5, size=(N, R, T)).astype(np.int32)
    actions = np.clip(rng.normal(-0.3, 0.6, size=(periods, N, m)), -1.2, 1.2)
    env.reset(customer_demand=demand)
    a_dev = torch.as_tensor(actions, device="cuda:0")
    rewards, variants = [], set()
    for t in range(periods):
        o, r, done, _ = env.step(a_dev[t])
        variants.add(env._lib.imx_kernel_variant(env._handle))
        rewards.append(torch.stack([r[n] for n in env.agent_names], dim=1) if env.MULTI else r[:, None])
    obs_last = (torch.stack([o[n] for n in env.agent_names], dim=1) if env.MULTI else o).cpu().numpy()
    st = {k: v.cpu().numpy() for k, v in env.state_dict().items()}
    assert int(env.error_flags.abs().sum()) == 0
    want = c_oracle.COracle(kind, cfg).run(demand, actions, periods=periods)
    return obs_last, torch.stack(rewards).cpu().numpy(), st, want, variants


@pytest.mark.parametrize("stages,ctas", [("2", "1"), ("3", "1"), ("4", "2"), ("4", "")])
@pytest.mark.parametrize("kind,preset,N", [("MAIM", "serial4", 6148), ("MAIM", "serial8", 5000), ("IM", "serial4", 4096 + 32),
                                           ("MAIM_div", "div1", 7172), ("MAIM_div", "div2", 9000), ("IM_div", "div2", 4100),
                                           ("MAIM", "serial2", 8192)])
def test_pipe_kernel_matches_c_oracle(kind, preset, N, stages, ctas, monkeypatch):
    monkeypatch.setenv("IMX_PIPE", "1")
    monkeypatch.setenv("IMX_PIPE_STAGES", stages)
    if ctas:
        monkeypatch.setenv("IMX_PIPE_CTAS", ctas)
    else:
        monkeypatch.delenv("IMX_PIPE_CTAS", raising=False)
    cfg = presets.PRESETS[preset]()
    rng = np.random.default_rng(hash((kind, preset, N)) % 2 ** 32)
    obs, rew, st, want, variants = _episode(kind, cfg, N, rng, periods=12)
    assert 3 in variants, variants                     # the pipelined kernel served the whole tiles
    np.testing.assert_array_equal(obs, want["obs_last"])
    np.testing.assert_array_equal(rew, want["reward"])
    for k in ("inv", "backlog", "order_u", "pipe"):
        np.testing.assert_array_equal(st[k], want[k], err_msg=k)
    if "backlog_to" in st:
        np.testing.assert_array_equal(st["backlog_to"], want["backlog_to"])


@pytest.mark.parametrize("hints,prefetch", [("0", "0"), ("1", "1"), ("2", "1"), ("1", "0")])
@pytest.mark.parametrize("kind,preset,N", [("MAIM", "serial4", 6148), ("MAIM_div", "div2", 9000), ("IM", "serial8", 3072 + 32)])
def test_l2_priorities_and_input_prefetch_do_not_change_results(kind, preset, N, hints, prefetch, monkeypatch):
    """The L2 eviction priorities on the bulk copies (IMX_L2_HINTS = 0 none / 1 all / 2 outputs only) and the L2 prefetch of the
    action tiles and demand rows ahead of griddepcontrol.wait (IMX_ACT_PREFETCH) are cache hints only: every combination gives the
    oracle's bits (the library picks them from the batch size; the switches force each variant here)."""
    monkeypatch.setenv("IMX_PIPE", "1")
    monkeypatch.setenv("IMX_L2_HINTS", hints)
    monkeypatch.setenv("IMX_ACT_PREFETCH", prefetch)
    cfg = presets.PRESETS[preset]()
    rng = np.random.default_rng(hash((kind, preset, N, hints, prefetch)) % 2 ** 32)
    obs, rew, st, want, variants = _episode(kind, cfg, N, rng, periods=10)
    assert 3 in variants, variants
    np.testing.assert_array_equal(obs, want["obs_last"])
    np.testing.assert_array_equal(rew, want["reward"])
    for k in ("inv", "backlog", "order_u", "pipe"):
        np.testing.assert_array_equal(st[k], want[k], err_msg=k)


def test_pipe_kernel_f32_obs_and_no_obs(monkeypatch):
    monkeypatch.setenv("IMX_PIPE", "1")
    monkeypatch.setenv("IMX_PIPE_CTAS", "1")
    cfg = presets.serial8()
    rng = np.random.default_rng(5)
    obs, rew, st, want, variants = _episode("MAIM", cfg, 4096, rng, periods=8, obs_dtype="float32")
    assert 3 in variants
    np.testing.assert_array_equal(obs, want["obs_last"].astype(np.float32))
    np.testing.assert_array_equal(rew, want["reward"])
    # obs = NULL through the raw ABI: rewards and state only
    import ctypes as C
    from marl_for_im_b200 import _lib
    N, T, m = 4096, 30, 8
    env = ENV_CLASSES["MAIM"](dict(cfg, num_envs=N))
    demand = rng.poisson(5, size=(N, 1, T)).astype(np.int32)
    actions = rng.uniform(-1, 1, size=(6, N, m))
    env.reset(customer_demand=demand)
    a = torch.as_tensor(actions, device="cuda:0")
    r = torch.empty((6, N, m), dtype=torch.float64, device="cuda:0")
    s = torch.cuda.current_stream().cuda_stream
    for t in range(6):
        _lib.check(env._lib.imx_step(env._handle, C.c_void_p(a[t].data_ptr()), None, C.c_void_p(r[t].data_ptr()), None, C.c_void_p(s)))
    assert env._lib.imx_kernel_variant(env._handle) == 3
    want = c_oracle.COracle("MAIM", cfg).run(demand, actions, periods=6)
    np.testing.assert_array_equal(r.cpu().numpy(), want["reward"])
    np.testing.assert_array_equal(env.state_dict()["inv"].cpu().numpy(), want["inv"])


def test_pipe_and_plain_kernels_agree_under_graph_replay(monkeypatch):
    """30 dependent launches captured in a CUDA graph and replayed (programmatic dependent launch between persistent grids)."""
    import ctypes as C
    from marl_for_im_b200 import _lib
    cfg = presets.serial4()
    N, T, m = 65536, 30, 4
    rng = np.random.default_rng(11)
    demand = torch.as_tensor(rng.poisson(5, size=(N, 1, T)).astype(np.int32), device="cuda:0")
    actions = torch.as_tensor(rng.uniform(-1, 1, size=(T, N, m)), device="cuda:0")
    outs = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("IMX_PIPE", mode)
        env = ENV_CLASSES["MAIM"](dict(cfg, num_envs=N))
        O = env.obs_len
        obs = torch.zeros((T, N, m, O), dtype=torch.float64, device="cuda:0")
        rew = torch.zeros((T, N, m), dtype=torch.float64, device="cuda:0")
        lib, h = env._lib, env._handle

        def run(stream):
            _lib.check(lib.imx_reset(h, C.c_void_p(demand.data_ptr()), None, 0, 1, None, C.c_void_p(stream)))
            for t in range(T):
                _lib.check(lib.imx_step(h, C.c_void_p(actions[t].data_ptr()), C.c_void_p(obs[t].data_ptr()), C.c_void_p(rew[t].data_ptr()),
                                        None, C.c_void_p(stream)))
        g = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            run(side.cuda_stream)
            side.synchronize()
            with torch.cuda.graph(g, stream=side):
                run(torch.cuda.current_stream().cuda_stream)
        torch.cuda.current_stream().wait_stream(side)
        assert lib.imx_kernel_variant(h) == (3 if mode == "1" else 2)
        for _ in range(5):
            obs.zero_(); rew.zero_()
            g.replay()
        torch.cuda.synchronize()
        outs[mode] = (obs.clone(), rew.clone(), env.state_dict()["pipe"].clone())
    for a, b in zip(outs["0"], outs["1"]):
        assert torch.equal(a, b)


@pytest.mark.parametrize("pipe", ["0", "1"])
@pytest.mark.parametrize("kind,preset,N", [("MAIM", "serial4", 4096), ("IM", "serial8", 3072 + 32), ("MAIM_div", "div2", 4096 + 64), ("IM_div", "div1", 2048)])
def test_noisy_delays_through_the_specialised_kernels(kind, preset, N, pipe, monkeypatch):
    """Episodes with noisy delays (replayed masks) are served by the runtime-specialised kernels too — one tile per CTA, pipelined,
    env-per-thread (div2) and multi-period — and match the C oracle on every env; an episode without noise on the same handle
    (the flag is per episode) still matches."""
    monkeypatch.setenv("IMX_PIPE", pipe)
    monkeypatch.setenv("IMX_PIPE_CTAS", "2")
    cfg = presets.PRESETS[preset]()
    rng = np.random.default_rng(N)
    env = ENV_CLASSES[kind](dict(cfg, num_envs=N))
    m, T, R = env.num_nodes, env.num_periods, len(env._retailers)
    co = c_oracle.COracle(kind, cfg)
    for noisy in (True, False, True):
        demand = rng.poisson(5, size=(N, R, T)).astype(np.int32)
        actions = np.clip(rng.normal(-0.3, 0.6, size=(T, N, m)), -1.2, 1.2)
        mask = (rng.uniform(size=(N, T, m)) <= 0.3) if noisy else None
        env.reset(customer_demand=demand, delay_mask=mask)
        a_dev = torch.as_tensor(actions, device="cuda:0")
        rews, variants = [], set()
        half = T // 2
        for t in range(half):
            o, r, done, _ = env.step(a_dev[t])
            variants.add(env._lib.imx_kernel_variant(env._handle))
            rews.append(torch.stack([r[n] for n in env.agent_names], dim=1) if env.MULTI else r[:, None])
        obs_many, rew_many, done = env.step_many(a_dev[half:])
        assert variants == {3 if pipe == "1" else 2}, variants
        want = co.run(demand, actions, mask)
        got = torch.cat([torch.stack(rews), rew_many.reshape(T - half, N, -1)]).cpu().numpy()
        np.testing.assert_array_equal(got, want["reward"])
        np.testing.assert_array_equal(obs_many[-1].cpu().numpy(), want["obs_last"])
        st = {k: v.cpu().numpy() for k, v in env.state_dict().items()}
        for k in ("inv", "backlog", "order_u", "pipe"):
            np.testing.assert_array_equal(st[k], want[k], err_msg=k)
