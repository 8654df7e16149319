"""Memory-safety canaries (compute-sanitizer is closed on this pool): every caller buffer sits between guard regions filled
with a sentinel, and the library's own 256-byte field padding is filled with the same sentinel after reset(); after
step / step_many / step_cc / rollout / host-buffer calls at awkward batch sizes (tile multiple +- 1, N % 4 != 0, tail tiles,
unaligned buffers -> direct kernel) every sentinel byte must be intact.  Bit-exact outputs alone would not catch a write
that lands in padding."""
import ctypes as C

import numpy as np
import pytest
import torch

from marl_for_im_b200 import _lib, presets
from marl_for_im_b200.envs import ENV_CLASSES, _DevView

pytestmark = pytest.mark.gpu
GUARD = 4096
SENT = 0xA5


class Guarded:
    """a device (or pinned host) buffer of `nbytes` with GUARD sentinel bytes on both sides; `ptr` is the interior pointer"""

    def __init__(self, nbytes, offset=0, pinned=False):
        self.nbytes, self.offset = int(nbytes), int(offset)
        total = 2 * GUARD + self.nbytes + 64
        self.raw = torch.full((total,), SENT, dtype=torch.uint8).pin_memory() if pinned else torch.full((total,), SENT, dtype=torch.uint8, device="cuda:0")
        base = self.raw.data_ptr()
        self.start = GUARD + ((-(base + GUARD)) % 256) + self.offset        # interior starts 256-byte aligned (+ a deliberate misalignment)
        self.ptr = base + self.start

    def view(self, dtype, shape):
        return self.raw[self.start:self.start + self.nbytes].view(dtype).reshape(shape)

    def check(self, what):
        r = self.raw.cpu().numpy() if self.raw.is_cuda else self.raw.numpy()
        assert (r[:self.start] == SENT).all(), f"{what}: write BEFORE the buffer"
        assert (r[self.start + self.nbytes:] == SENT).all(), f"{what}: write BEHIND the buffer"


def _pad_views(env):
    """uint8 views of the library-owned padding behind every state field (fields are 256-byte aligned inside one block)"""
    out = []
    for f in range(8):                                   # IMX_F_INV .. IMX_F_BACKLOG_TO
        ptr, cnt = C.c_void_p(), C.c_int64()
        _lib.check(env._lib.imx_state_field(env._handle, f, C.byref(ptr), C.byref(cnt)))
        used = cnt.value * 4
        pad = (-used) % 256
        if cnt.value and pad:
            out.append((f, torch.as_tensor(_DevView(ptr.value + used, (pad,), "|u1"), device="cuda:0")))
    return out


@pytest.mark.parametrize("kind,preset", [("MAIM", "serial4"), ("IM", "serial8"), ("MAIM_div", "div2"), ("IM_div", "div1"), ("MAIM", "serial2")])
@pytest.mark.parametrize("N", [4095, 4097, 4098, 6148, 8192, 33])
@pytest.mark.parametrize("misalign", [0, 8])
def test_no_write_outside_buffers(kind, preset, N, misalign, monkeypatch):
    if N >= 4096:
        monkeypatch.setenv("IMX_PIPE_CTAS", "2")        # the pipelined kernel walks several tiles per CTA at these sizes
    cfg = presets.PRESETS[preset]()
    noisy = kind in ("MAIM", "IM_div") and N in (4097, 33)
    env = ENV_CLASSES[kind](dict(cfg, num_envs=N))
    if noisy:
        env.reset(noisy_delay=True, noisy_delay_threshold=0.3, customer_demand=np.full((len(env._retailers), 30), 5))
    m, T, O, R = env.num_nodes, env.num_periods, env.obs_len, len(env._retailers)
    cols = m if env.MULTI else 1
    W = (m - 1) * (1 + O) + O
    K = 5
    rng = np.random.default_rng(N)
    g_dem = Guarded(N * R * T * 4)
    g_act = Guarded(K * N * m * 8, misalign)
    g_obs = Guarded(K * N * m * O * 8, misalign)
    g_rew = Guarded(K * N * cols * 8, misalign)
    g_cc = Guarded(N * m * W * 8)
    g_z = Guarded(N * m * 8)
    g_ret = Guarded(N * cols * 8)
    g_sr = Guarded(T * N * cols * 8)
    g_pmf = Guarded(N * R * T * 8)
    g_dfo = Guarded(N * 8)
    g_mask = Guarded(N * T * m)
    everything = dict(dem=g_dem, act=g_act, obs=g_obs, rew=g_rew, cc=g_cc, z=g_z, ret=g_ret, sr=g_sr, pmf=g_pmf, dfo=g_dfo, mask=g_mask)
    g_dem.view(torch.int32, (N, R, T)).copy_(torch.as_tensor(rng.poisson(5, size=(N, R, T)).astype(np.int32)))
    g_act.view(torch.float64, (K, N, m)).copy_(torch.as_tensor(rng.uniform(-1.2, 1.2, size=(K, N, m))))
    g_z.view(torch.float64, (N, m)).copy_(torch.as_tensor(rng.integers(5, 40, size=(N, m)).astype(np.float64)))
    g_pmf.view(torch.float64, (N, R, T)).copy_(torch.as_tensor(rng.uniform(0, 0.2, size=(N, R, T))))
    g_mask.view(torch.uint8, (N, T, m)).copy_(torch.as_tensor((rng.uniform(size=(N, T, m)) < 0.3).astype(np.uint8)))
    lib, h = env._lib, env._handle
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    P = C.c_void_p

    def check_all(what):
        torch.cuda.synchronize()
        for name, g in everything.items():
            g.check(f"{what}: {name}")
        for f, v in pads:
            assert bool((v == SENT).all()), f"{what}: library padding behind state field {f} was overwritten"

    _lib.check(lib.imx_reset(h, P(g_dem.ptr), P(g_mask.ptr) if noisy else None, int(noisy), 1, P(g_obs.ptr), s))
    torch.cuda.synchronize()
    pads = _pad_views(env)
    for _, v in pads:
        v.fill_(SENT)
    check_all("reset")
    _lib.check(lib.imx_step(h, P(g_act.ptr), P(g_obs.ptr), P(g_rew.ptr), None, s))
    check_all("step")
    _lib.check(lib.imx_step(h, P(g_act.ptr), None, P(g_rew.ptr), None, s))
    check_all("step without observations")
    if not noisy:
        _lib.check(lib.imx_step_many(h, P(g_act.ptr), K, P(g_obs.ptr), P(g_rew.ptr), None, s))
        check_all("step_many")
    if env.MULTI and not misalign:
        _lib.check(lib.imx_step_cc(h, P(g_act.ptr), P(g_obs.ptr), P(g_cc.ptr), 1, -1.0, 1.0, P(g_rew.ptr), s))
        check_all("step_cc")
    has_dist = env.demand_dist in ("poisson", "uniform")
    _lib.check(lib.imx_rollout_basestock(h, P(g_z.ptr), m, P(g_dem.ptr), P(g_mask.ptr) if noisy else None, int(noisy), 2,
                                         None if env.MULTI else P(g_pmf.ptr), P(g_ret.ptr), P(g_sr.ptr), None if env.MULTI else P(g_dfo.ptr), 1, s))
    check_all("rollout (replayed demand, write_state)")
    if has_dist:
        _lib.check(lib.imx_rollout_basestock(h, P(g_z.ptr), 0, None, None, int(noisy), 3, None, P(g_ret.ptr), None, None, 0, s))
        check_all("rollout (Philox demand)")
    # host-buffer path on guarded pinned memory (zero-copy: the kernels address these host bytes directly)
    h_act, h_obs, h_rew = Guarded(N * m * 8, pinned=True), Guarded(N * m * O * 8, pinned=True), Guarded(N * cols * 8, pinned=True)
    h_dem = Guarded(N * R * T * 4, pinned=True)
    h_dem.view(torch.int32, (N, R, T)).copy_(g_dem.view(torch.int32, (N, R, T)).cpu())
    h_act.view(torch.float64, (N, m)).copy_(g_act.view(torch.float64, (K, N, m))[0].cpu())
    _lib.check(lib.imx_reset_host(h, P(h_dem.ptr), None, 0, 5, P(h_obs.ptr)))
    for _, v in pads:
        v.fill_(SENT)
    _lib.check(lib.imx_step_host(h, P(h_act.ptr), P(h_obs.ptr), P(h_rew.ptr)))
    check_all("host-buffer step")
    for name, g in dict(h_act=h_act, h_obs=h_obs, h_rew=h_rew, h_dem=h_dem).items():
        g.check(f"host-buffer step: {name}")
