"""no entry point may compile, load a module or allocate on the step path.  The specialised kernels are loaded
by imx_create(); imx_prepare() loads the remaining variants and scratch buffers; the FIRST step of a handle can be captured in a
CUDA graph without any warm-up call, and a capture in progress never triggers a compile."""
import ctypes as C

import numpy as np
import pytest
import torch

from marl_for_im_b200 import _lib, presets
from marl_for_im_b200.envs import ENV_CLASSES
from oracle import c_oracle

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind,preset,with_obs,prepare", [("MAIM", "serial4", True, False), ("MAIM", "serial4", False, True), ("MAIM_div", "div2", True, True),
                                                          ("MAIM", "serial2", False, False)])
def test_first_step_is_capturable_without_warmup(kind, preset, with_obs, prepare):
    cfg = presets.PRESETS[preset]()
    N, K = 8192, 6
    env = ENV_CLASSES[kind](dict(cfg, num_envs=N))
    lib, h = env._lib, env._handle
    if prepare:
        _lib.check(lib.imx_prepare(h, 0))
    m, T, O, R = env.num_nodes, env.num_periods, env.obs_len, len(env._retailers)
    rng = np.random.default_rng(1)
    demand_h = rng.poisson(5, size=(N, R, T)).astype(np.int32)
    actions_h = rng.uniform(-1, 1, size=(K, N, m))
    demand, actions = torch.as_tensor(demand_h, device="cuda:0"), torch.as_tensor(actions_h, device="cuda:0")
    obs = torch.zeros((K, N, m, O), dtype=torch.float64, device="cuda:0")
    rew = torch.zeros((K, N, m), dtype=torch.float64, device="cuda:0")
    launches0 = env.launch_count()
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):            # NO warm-up call before the capture
            s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
            _lib.check(lib.imx_reset(h, C.c_void_p(demand.data_ptr()), None, 0, 1, None, s))
            for t in range(K):
                _lib.check(lib.imx_step(h, C.c_void_p(actions[t].data_ptr()), C.c_void_p(obs[t].data_ptr()) if with_obs else None,
                                        C.c_void_p(rew[t].data_ptr()), None, s))
    torch.cuda.current_stream().wait_stream(side)
    assert env.launch_count() - launches0 >= K + 1
    if with_obs or prepare:                                # without imx_prepare the obs-less variant is not loaded under capture: AOT kernel
        assert lib.imx_kernel_variant(h) in (2, 3)
    else:
        assert lib.imx_kernel_variant(h) in (0, 1)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    want = c_oracle.COracle(kind, cfg).run(demand_h, actions_h, periods=K, all_obs=True)
    np.testing.assert_array_equal(rew.cpu().numpy(), want["reward"])
    if with_obs:
        np.testing.assert_array_equal(obs.cpu().numpy(), want["obs_all"][1:])
    np.testing.assert_array_equal(env.state_dict()["inv"].cpu().numpy(), want["inv"])


def test_prepare_flags_and_dfo_scratch():
    cfg = presets.serial4_dfo()
    env = ENV_CLASSES["IM"](dict(cfg, num_envs=2048))
    lib, h = env._lib, env._handle
    _lib.check(lib.imx_prepare(h, 8 | 4))                 # IMX_PREPARE_DFO | IMX_PREPARE_HOST
    assert lib.imx_prepare(None, 0) < 0
    # critic rows are a multi-agent notion: the flag is ignored for a single-agent handle, the call still succeeds
    _lib.check(lib.imx_prepare(h, 16))
