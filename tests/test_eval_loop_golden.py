"""Evaluation-loop accumulators: the oracle restatement (CPU) and the CUDA accumulators (GPU) against the fixture
generated from the reference by tests/golden/make_golden_eval.py."""
import json
import os

import numpy as np
import pytest

from harness import GOLDEN_DIR, copy_config
from marl_for_im_b200 import presets
from oracle import im_oracle


def load_cases():
    z = np.load(os.path.join(GOLDEN_DIR, "evalloop", "eval_loop.npz"), allow_pickle=False)
    cases = json.loads(str(z["cases"]))
    return [(k, p, kw, r, z[f"demand_{i}"], z[f"actions_{i}"], z[f"want_{i}"]) for i, (k, p, kw, r) in enumerate(cases)]


def case_config(preset, kw):
    cfg = presets.PRESETS[preset](**kw)
    m = cfg.get("num_nodes", cfg.get("num_stages"))
    if preset != "serial4_dfo":
        cfg["inv_max"] = np.array([30, 25, 40, 35, 30, 45, 20, 30][:m], dtype=float)
    return cfg


CASES = load_cases()
IDS = [f"{c[0]}-{c[1]}" for c in CASES]


@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_oracle_eval_loop_matches_reference_fixture(case):
    kind, preset, kw, rescaled, demand, actions, want = case
    cfg = case_config(preset, kw)
    for e in range(demand.shape[0]):
        got = im_oracle.eval_loop_accumulators(im_oracle.OracleEnv(kind, copy_config(cfg)), demand[e], actions[e], rescaled)
        np.testing.assert_array_equal(got, want[e])


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_cuda_eval_accumulators_match_reference_fixture(case):
    """Each fixture episode replicated over a batch with distinct extra envs; accumulators bit-exact."""
    import torch
    from marl_for_im_b200.envs import ENV_CLASSES
    kind, preset, kw, rescaled, demand, actions, want = case
    cfg = case_config(preset, kw)
    E = demand.shape[0]
    reps = 50                                                     # N = 300: several CTAs, tail tile
    N = E * reps
    dem = np.repeat(demand, reps, axis=0).astype(np.int32)
    act = np.ascontiguousarray(np.repeat(actions, reps, axis=0).transpose(1, 0, 2))     # [T, N, m]
    env = ENV_CLASSES[kind](dict(copy_config(cfg), num_envs=N, return_info=True))
    env.reset(customer_demand=dem)
    a_dev = torch.as_tensor(act, device="cuda:0")
    acc = None
    for t in range(env.num_periods):
        o, r, done, info = env.step(a_dev[t])
        profit = torch.stack([info[a]["profit"] for a in env.agent_names], dim=1) if env.MULTI else info["profit"]
        acc = env.eval_accumulate(acc, o, r, profit)
    got = acc.cpu().numpy()
    np.testing.assert_array_equal(got, np.repeat(want, reps, axis=0))
    # the column statistics: {n, (Σ, Σ²) per column}, accumulate across evaluation batches
    st = env.eval_stats(acc).cpu().numpy()
    assert st[0] == N and st.shape == (1 + 2 * got.shape[1],)
    np.testing.assert_allclose(st[1::2], got.sum(axis=0), rtol=1e-12, atol=1e-9)
    np.testing.assert_allclose(st[2::2], (got * got).sum(axis=0), rtol=1e-12, atol=1e-9)
    st2 = env.eval_stats(acc, stats=torch.as_tensor(st, device="cuda:0").clone(), accumulate=True).cpu().numpy()
    np.testing.assert_array_equal(st2, 2 * st)
    summ = env.eval_summary(st, env.num_nodes)
    np.testing.assert_allclose(summ["episode_reward"][0], np.mean(got[:, 0]), rtol=1e-12)
    np.testing.assert_allclose(summ["total_inventory"][1], np.std(got[:, 1]), rtol=1e-9, atol=1e-9)
    # a second episode with reset=True starts from zero again and gives the same rows (profit omitted: columns stay 0)
    env.reset(customer_demand=dem)
    acc2 = acc.clone()
    for t in range(env.num_periods):
        o, r, done, info = env.step(a_dev[t])
        acc2 = env.eval_accumulate(acc2, o, r, None, reset=(t == 0))
    got2 = acc2.cpu().numpy()
    np.testing.assert_array_equal(got2[:, :4], got[:, :4])
    assert not got2[:, 4:].any()


@pytest.mark.gpu
def test_cuda_eval_accumulators_on_float32_observations():
    """An env that writes float32 observations: the accumulators undo the scaling on the float32 values widened to float64
    (what a script reading RLlib's float32 observations would compute), operation for operation."""
    import torch
    from marl_for_im_b200.envs import ENV_CLASSES
    cfg = case_config("serial4", {})
    N, m = 512, 4
    rng = np.random.default_rng(32)
    demand = rng.poisson(6, size=(N, 30)).astype(np.int32)
    actions = torch.as_tensor(rng.uniform(-1, 1, size=(30, N, m)), device="cuda:0")
    env = ENV_CLASSES["MAIM"](dict(copy_config(cfg), num_envs=N, obs_dtype="float32"))
    env.reset(customer_demand=demand)
    acc = None
    want = torch.zeros((N, 4), dtype=torch.float64, device="cuda:0")
    inv_max = torch.as_tensor(np.asarray(cfg["inv_max"], dtype=np.float64), device="cuda:0")
    a, b = float(cfg["a"]), float(cfg["b"])
    for t in range(30):
        o, r, _, _ = env.step(actions[t])
        acc = env.eval_accumulate(acc, o, r)
        x = env.last_obs.double()                                   # [N, m, O]
        undo = lambda col: (((x[:, :, col] - a) * (inv_max - 0.0)) / (b - a)) + 0.0     # noqa: E731  rev_scale, MAIM_env.py:509-519
        step_inv = torch.zeros(N, dtype=torch.float64, device="cuda:0")
        step_bl = torch.zeros(N, dtype=torch.float64, device="cuda:0")
        for i in range(m):
            want[:, 0] += env.last_reward[:, i]
            step_inv += undo(0)[:, i]
            step_bl += undo(1)[:, i]
        want[:, 1] += step_inv
        want[:, 2] += step_bl
        want[:, 3] += undo(1)[:, 0]
    assert torch.equal(acc[:, :4], want)
