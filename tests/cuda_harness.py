"""Runs one episode through the CUDA path (batched env classes → ctypes → libimx_b200.so) and
returns the same dict of arrays as harness.run_oracle / run_reference."""
import numpy as np
import torch


def run_cuda(kind, cfg, demand, actions, delay_mask=None, n_copies=1):
    """Runs one episode on the GPU in batched mode (the same trace replicated ``n_copies`` times)
    and returns the harness dict for env 0 plus the raw batched tensors."""
    from marl_for_im_b200.envs import ENV_CLASSES
    from harness import copy_config
    c = copy_config(cfg)
    c.update(num_envs=n_copies, device="cuda:0", return_info=True)
    env = ENV_CLASSES[kind](c)
    m, T = env.num_nodes, env.num_periods
    multi = kind.startswith("MAIM")
    dm = None
    if delay_mask is not None:
        dm = np.broadcast_to(np.asarray(delay_mask)[None], (n_copies,) + np.asarray(delay_mask).shape)
    d = np.asarray(demand)
    d = np.broadcast_to(d[None], (n_copies,) + d.shape)
    o = env.reset(customer_demand=d, delay_mask=dm)

    def pack(o):
        return (torch.stack([o[n] for n in env.agent_names], dim=1) if multi else o).cpu().numpy()

    obs = [pack(o)]
    out = {k: np.zeros((T, n_copies, m)) for k in ("reward", "demand", "ship", "acq", "order", "profit")}
    st = {k: [env.state_dict()[k].cpu().numpy().astype(np.int64)] for k in ("inv", "backlog", "order_u")}
    for t in range(T):
        a = torch.as_tensor(np.broadcast_to(actions[t][None], (n_copies, m)).copy(), device="cuda:0")
        if multi:
            a = {n: a[:, i] for i, n in enumerate(env.agent_names)}
        o, r, done, info = env.step(a)
        obs.append(pack(o))
        if multi:
            out["reward"][t] = torch.stack([r[n] for n in env.agent_names], dim=1).cpu().numpy()
            for i, n in enumerate(env.agent_names):
                out["demand"][t, :, i] = info[n]["demand"].cpu().numpy()
                out["ship"][t, :, i] = info[n]["ship"].cpu().numpy()
                out["acq"][t, :, i] = info[n]["acquisition"].cpu().numpy()
                out["order"][t, :, i] = info[n]["actual order"].cpu().numpy()
                out["profit"][t, :, i] = info[n]["profit"].cpu().numpy()
            assert done["__all__"] == (t == T - 1)
        else:
            out["reward"][t, :, 0] = r.cpu().numpy()
            out["demand"][t] = info["demand"].cpu().numpy()
            out["ship"][t] = info["ship"].cpu().numpy()
            out["acq"][t] = info["acquisition"].cpu().numpy()
            out["order"][t] = info["actual order"].cpu().numpy()
            out["profit"][t] = info["profit"].cpu().numpy()
            assert done == (t == T - 1)
        for k in st:
            st[k].append(env.state_dict()[k].cpu().numpy().astype(np.int64))
    assert int(env.error_flags.abs().sum()) == 0
    res = {k: v[:, 0] for k, v in out.items()}
    res["obs"] = np.stack(obs)[:, 0]
    for k in st:
        res[k] = np.stack(st[k])[:, 0]
    # every replica must agree with env 0
    full_obs = np.stack(obs)
    assert np.array_equal(full_obs, np.broadcast_to(full_obs[:, :1], full_obs.shape))
    return res


def assert_same(a, b, what=""):
    for k in ("inv", "backlog", "order_u", "demand", "ship", "acq", "order", "profit", "reward", "obs"):
        np.testing.assert_array_equal(a[k], b[k], err_msg=f"{what}: {k}")
