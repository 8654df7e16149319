"""imx_step_many (K periods on a stored plan in one launch, state resident in shared memory, double-buffered inputs and
outputs) against K plain step() calls: same bytes for observations, rewards and final state — at sizes that use the
ahead-of-time kernel, the runtime-specialised kernel, a tail tile (falls back to plain launches), split calls, under
CUDA-graph replay, and repeated many times at the benchmark size (a buffer-reuse bug would show up as a data race)."""
import ctypes as C

import numpy as np
import pytest
import torch

from harness import copy_config
from marl_for_im_b200 import _lib, presets
from marl_for_im_b200.envs import ENV_CLASSES

pytestmark = pytest.mark.gpu


def _plain(kind, cfg, demand, actions, n):
    env = ENV_CLASSES[kind](dict(copy_config(cfg), num_envs=n))
    env.reset(customer_demand=demand)
    obs, rew = [], []
    for t in range(actions.shape[0]):
        env.step(actions[t])
        obs.append(env.last_obs.clone())
        rew.append(env.last_reward.clone())
    return torch.stack(obs), torch.stack(rew), {k: v.clone() for k, v in env.state_dict().items()}


def _inputs(kind, cfg, n, seed):
    env = ENV_CLASSES[kind](dict(copy_config(cfg), num_envs=4))
    m, T, R = env.num_nodes, env.num_periods, len(env._retailers)
    g = torch.Generator(device="cuda:0")
    g.manual_seed(seed)
    demand = torch.poisson(torch.full((n, R, T), 5.0, device="cuda:0"), generator=g).to(torch.int32)
    if kind.endswith("div"):
        actions = (torch.randn((T, n, m), dtype=torch.float64, device="cuda:0", generator=g) * 0.5 - 0.6).clamp(-1, 1)
    else:
        actions = torch.rand((T, n, m), dtype=torch.float64, device="cuda:0", generator=g) * 2.3 - 1.15
    return demand, actions


@pytest.mark.parametrize("kind,preset", [("MAIM", "serial4"), ("IM", "serial8"), ("MAIM_div", "div2"), ("IM_div", "div1")])
@pytest.mark.parametrize("n", [96, 100, 4096, 4096 + 36, 65536])
def test_step_many_equals_plain_steps(kind, preset, n):
    cfg = presets.PRESETS[preset]()
    demand, actions = _inputs(kind, cfg, n, seed=n)
    want_obs, want_rew, want_state = _plain(kind, cfg, demand, actions, n)
    env = ENV_CLASSES[kind](dict(copy_config(cfg), num_envs=n))
    T = env.num_periods
    for split in ((T,), (1, T - 1), (7, 11, T - 18)):
        env.reset(customer_demand=demand)
        t0, obs, rew = 0, [], []
        for k in split:
            o, r, done = env.step_many(actions[t0:t0 + k])
            obs.append(o)
            rew.append(r)
            t0 += k
            assert (done["__all__"] if env.MULTI else done) == (t0 == T)
        assert torch.equal(torch.cat(obs), want_obs) and torch.equal(torch.cat(rew), want_rew), (kind, n, split)
        for k, v in env.state_dict().items():
            assert torch.equal(v, want_state[k]), (kind, n, split, k)
    with pytest.raises(IndexError):
        env.step_many(actions[:1])


def test_step_many_repeated_and_graph_replayed_at_bench_size():
    kind, cfg, n = "MAIM", presets.serial4(), 65536
    demand, actions = _inputs(kind, cfg, n, seed=1)
    want_obs, want_rew, want_state = _plain(kind, cfg, demand, actions, n)
    env = ENV_CLASSES[kind](dict(copy_config(cfg), num_envs=n))
    T = env.num_periods
    obs = torch.empty_like(want_obs)
    rew = torch.empty_like(want_rew)
    for rep in range(25):
        env.reset(customer_demand=demand)
        obs.zero_()
        env.step_many(actions, obs_out=obs, reward_out=rew)
        assert torch.equal(obs, want_obs) and torch.equal(rew, want_rew), rep
    assert env._lib.imx_kernel_variant(env._handle) in (2, 3)
    # one CUDA graph: reset + the chained 30-period replay; replayed with fresh output buffers each time
    lib, h = env._lib, env._handle
    side = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()

    def episode(stream):
        _lib.check(lib.imx_reset(h, C.c_void_p(demand.data_ptr()), None, 0, 1, None, C.c_void_p(stream)))
        _lib.check(lib.imx_step_many(h, C.c_void_p(actions.data_ptr()), T, C.c_void_p(obs.data_ptr()), C.c_void_p(rew.data_ptr()),
                                     None, C.c_void_p(stream)))
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        episode(side.cuda_stream)
        side.synchronize()
        with torch.cuda.graph(g, stream=side):
            episode(torch.cuda.current_stream().cuda_stream)
    torch.cuda.current_stream().wait_stream(side)
    for rep in range(25):
        obs.zero_()
        rew.zero_()
        g.replay()
        g.replay()                      # back-to-back replays
        torch.cuda.synchronize()
        assert torch.equal(obs, want_obs) and torch.equal(rew, want_rew), rep
    for k, v in env.state_dict().items():
        assert torch.equal(v, want_state[k]), k


@pytest.mark.parametrize("kind,preset", [("MAIM", "serial8"), ("MAIM_div", "div2")])
def test_step_many_with_noisy_delay_mask(kind, preset):
    """The carry state of a noisy lead time lives in the tile like the rest of the state; the replayed mask is read per period."""
    cfg = presets.PRESETS[preset]()
    n = 4096
    demand, actions = _inputs(kind, cfg, n, seed=5)
    env_a = ENV_CLASSES[kind](dict(copy_config(cfg), num_envs=n))
    env_b = ENV_CLASSES[kind](dict(copy_config(cfg), num_envs=n))
    T, m = env_a.num_periods, env_a.num_nodes
    mask = torch.rand((n, T, m), device="cuda:0") < 0.3
    env_a.reset(customer_demand=demand, delay_mask=mask)
    obs, rew = [], []
    for t in range(T):
        env_a.step(actions[t])
        obs.append(env_a.last_obs.clone())
        rew.append(env_a.last_reward.clone())
    env_b.reset(customer_demand=demand, delay_mask=mask)
    o, r, _ = env_b.step_many(actions)
    assert torch.equal(o, torch.stack(obs)) and torch.equal(r, torch.stack(rew))
    for k, v in env_b.state_dict().items():
        assert torch.equal(v, env_a.state_dict()[k]), k
    assert int(env_b.state_dict()["carry"].abs().sum()) >= 0 and "carry" in env_b.state_dict()


@pytest.mark.parametrize("kind,preset,n", [("MAIM", "serial4", 8192), ("IM_div", "div2", 4096), ("MAIM", "serial8", 100)])
def test_without_observation_output(kind, preset, n):
    """obs = NULL (a caller that only wants rewards, e.g. scoring a stored plan): same rewards and final state as with
    observations, through imx_step and through imx_step_many."""
    import ctypes as C
    from marl_for_im_b200 import _lib
    cfg = presets.PRESETS[preset]()
    demand, actions = _inputs(kind, cfg, n, seed=3)
    want_obs, want_rew, want_state = _plain(kind, cfg, demand, actions, n)
    env = ENV_CLASSES[kind](dict(copy_config(cfg), num_envs=n))
    T = env.num_periods
    env.reset(customer_demand=demand)
    o, r, _ = env.step_many(actions, want_obs=False)
    assert o is None and torch.equal(r, want_rew)
    for k, v in env.state_dict().items():
        assert torch.equal(v, want_state[k]), k
    env.reset(customer_demand=demand)
    rew = torch.empty_like(want_rew)
    for t in range(T):
        _lib.check(env._lib.imx_step(env._handle, C.c_void_p(actions[t].data_ptr()), None, C.c_void_p(rew[t].data_ptr()), None,
                                     C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    assert torch.equal(rew, want_rew)
    for k, v in env.state_dict().items():
        assert torch.equal(v, want_state[k]), k


@pytest.mark.parametrize("kind,preset,n", [("IM", "serial4", 4096), ("MAIM_div", "div2", 4096), ("MAIM", "serial8", 100)])
def test_step_many_with_diagnostics(kind, preset, n):
    """return_info=True: the per-period demand / ship / acquisition / order / profit blocks equal what K step() calls report."""
    cfg = presets.PRESETS[preset]()
    demand, actions = _inputs(kind, cfg, n, seed=9)
    ref = ENV_CLASSES[kind](dict(copy_config(cfg), num_envs=n, return_info=True))
    ref.reset(customer_demand=demand)
    want = {k: [] for k in ("demand", "ship", "acquisition", "actual order", "profit")}
    rew = []
    for t in range(ref.num_periods):
        _, r, _, info = ref.step(actions[t])
        for k in want:
            want[k].append(torch.stack([info[a][k] for a in ref.agent_names], dim=1) if ref.MULTI else info[k])
        rew.append(ref.last_reward.clone())
    env = ENV_CLASSES[kind](dict(copy_config(cfg), num_envs=n))
    env.reset(customer_demand=demand)
    o, r, done, info = env.step_many(actions, return_info=True)
    assert torch.equal(r, torch.stack(rew))
    for k in want:
        assert torch.equal(info[k], torch.stack(want[k]).to(info[k].dtype)), k
