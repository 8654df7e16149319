"""float32 observation output (``obs_dtype="float32"`` / ``imx_config.obs_f32``): the kernels compute the
reference's float64 observation and round it once to float32 — exactly ``np.float32(obs64)``, the cast RLlib's
preprocessor applies to every observation.  Rewards and integer state do not change."""
import ctypes as C

import numpy as np
import pytest
import torch

from harness import copy_config, random_case
from marl_for_im_b200 import _lib, presets
from marl_for_im_b200.envs import ENV_CLASSES

pytestmark = pytest.mark.gpu


def _episode(kind, cfg, demand, actions, n, obs_dtype):
    c = copy_config(cfg)
    c.update(num_envs=n, return_info=False, obs_dtype=obs_dtype)
    env = ENV_CLASSES[kind](c)
    multi = kind.startswith("MAIM")
    pack = lambda o: (torch.stack([o[a] for a in env.agent_names], dim=1) if multi else o).cpu().numpy()   # noqa: E731
    obs = [pack(env.reset(customer_demand=demand))]
    rew = []
    a_dev = torch.as_tensor(actions, device="cuda:0")
    for t in range(env.num_periods):
        o, r, _, _ = env.step(a_dev[t])
        obs.append(pack(o))
        rew.append((torch.stack([r[a] for a in env.agent_names], dim=1) if multi else r).cpu().numpy())
    st = {k: v.cpu().numpy() for k, v in env.state_dict().items()}
    return np.stack(obs), np.stack(rew), st, env._lib.imx_kernel_variant(env._handle)


CASES = [("MAIM", "serial4", {}), ("MAIM", "serial8", dict(prev_actions=True, prev_length=3, independent=True)),
         ("IM", "serial4", dict(prev_actions=True, prev_length=2)), ("MAIM", "serial2", {}),
         ("IM", "serial4_dfo", {}), ("MAIM", "serial4", dict(standardise_state=False, standardise_actions=False)),
         ("MAIM_div", "div1", {}), ("MAIM_div", "div2", dict(share_network=True, prev_actions=True)),
         ("IM_div", "div2", dict(prev_length=2))]


@pytest.mark.parametrize("n", [96, 97, 4096 + 8])
@pytest.mark.parametrize("kind,preset,kw", CASES)
def test_f32_obs_is_the_cast_of_f64_obs(kind, preset, kw, n, monkeypatch):
    if n == 96:
        monkeypatch.setenv("IMX_JIT", "0")           # ahead-of-time TMA kernel
    rng = np.random.default_rng(31 + n)
    cfg = presets.PRESETS[preset](**kw)
    T = cfg["num_periods"]
    demand, actions = [], []
    for _ in range(n):
        d, a = random_case(kind, cfg, rng, mu=6, action_mode="near_eq" if kind.endswith("div") else "uniform")
        demand.append(d)
        actions.append(a)
    demand = np.stack(demand).astype(np.int32)
    actions = np.ascontiguousarray(np.stack(actions, axis=1))          # [T, n, m]
    o64, r64, s64, _ = _episode(kind, cfg, demand, actions, n, "float64")
    o32, r32, s32, variant = _episode(kind, cfg, demand, actions, n, "float32")
    assert o32.dtype == np.float32 and o64.dtype == np.float64 and o32.shape == o64.shape and o32.shape[0] == T + 1
    np.testing.assert_array_equal(o32, o64.astype(np.float32))
    np.testing.assert_array_equal(r32, r64)
    for k in s64:
        np.testing.assert_array_equal(s32[k], s64[k], err_msg=k)
    if n >= 4096 and n % 4 == 0:
        assert variant in (2, 3)


def test_f32_obs_host_buffer_path(monkeypatch):
    cfg = presets.serial4()
    N, T, m = 4096 + 64, 30, 4
    rng = np.random.default_rng(8)
    demand = rng.poisson(5, size=(N, 1, T)).astype(np.int32)
    actions = rng.uniform(-1, 1, size=(T, N, m))
    o64, r64, _, _ = _episode("MAIM", cfg, demand, actions, N, "float64")
    for zero_copy in ("1", "0"):
        monkeypatch.setenv("IMX_HOST_ZERO_COPY", zero_copy)
        env = ENV_CLASSES["MAIM"](dict(copy_config(cfg), num_envs=N, obs_dtype="float32"))
        O = env.obs_len
        dem_h, act_h = torch.as_tensor(demand).pin_memory(), torch.as_tensor(actions).pin_memory()
        obs_h = torch.empty((N, m, O), dtype=torch.float32).pin_memory()
        rew_h = torch.empty((N, m), dtype=torch.float64).pin_memory()
        p = lambda a: C.c_void_p(a.data_ptr())   # noqa: E731
        _lib.check(env._lib.imx_reset_host(env._handle, p(dem_h), None, 0, 3, p(obs_h)))
        np.testing.assert_array_equal(obs_h.numpy(), o64[0].astype(np.float32))
        for t in range(T):
            _lib.check(env._lib.imx_step_host(env._handle, p(act_h[t]), p(obs_h), p(rew_h)))
            np.testing.assert_array_equal(obs_h.numpy(), o64[t + 1].astype(np.float32), err_msg=f"t={t}")
            np.testing.assert_array_equal(rew_h.numpy(), r64[t])
