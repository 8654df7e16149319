"""Vector adapters (marl_for_im_b200/vector.py): the BaseEnv / VectorEnv shaped surfaces over one GPU batch give,
env by env, what N separate reference-style envs give (checked against the oracle on the batch's own demand)."""
import numpy as np
import pytest
import torch

from harness import copy_config, run_oracle
from marl_for_im_b200 import presets
from marl_for_im_b200.envs import ENV_CLASSES
from marl_for_im_b200.vector import BatchedMultiAgentEnv, BatchedVectorEnv

pytestmark = pytest.mark.gpu


def _traces(env, kind):
    d = env.customer_demand_device().cpu().numpy()                # [T, R, N]
    return [d[:, :, n].T if kind.endswith("div") else d[:, 0, n] for n in range(env.num_envs)]


@pytest.mark.parametrize("kind,preset", [("MAIM", "serial4"), ("MAIM_div", "div2")])
def test_base_env_surface_two_episodes(kind, preset):
    cfg = presets.PRESETS[preset]()
    N = 7
    env = ENV_CLASSES[kind](dict(copy_config(cfg), num_envs=N, seed=3))
    venv = BatchedMultiAgentEnv(env)
    m, T = env.num_nodes, env.num_periods
    agents = env.agent_names
    rng = np.random.default_rng(0)
    for episode in range(2):
        actions = rng.uniform(-1, 1, size=(T, N, m))
        want = [run_oracle(kind, cfg, tr, actions[:, n]) for n, tr in enumerate(_traces(env, kind))]
        if episode == 0:
            obs, rew, dones, infos, off = venv.poll()
            assert all(rew[n] == {} for n in range(N)) and off == {}
        else:
            obs = {n: venv.try_reset(n) for n in range(N)}
            # the batch was reset exactly once: a new demand trace per env, drawn for the new episode
            want = [run_oracle(kind, cfg, tr, actions[:, n]) for n, tr in enumerate(_traces(env, kind))]
        for n in range(N):
            for i, a in enumerate(agents):
                np.testing.assert_array_equal(obs[n][a], want[n]["obs"][0, i])
        for t in range(T):
            venv.send_actions({n: {a: np.array([actions[t, n, i]]) for i, a in enumerate(agents)} for n in range(N)})
            obs, rew, dones, infos, _ = venv.poll()
            assert venv.poll()[0] == {}                           # nothing new until the next send_actions
            for n in range(N):
                assert dones[n]["__all__"] == (t == T - 1)
                for i, a in enumerate(agents):
                    np.testing.assert_array_equal(obs[n][a], want[n]["obs"][t + 1, i])
                    assert rew[n][a] == want[n]["reward"][t, i]
    with pytest.raises(ValueError):
        venv.send_actions({0: {a: 0.0 for a in agents}})
    # tensor surface: same numbers, no host objects
    venv.reset_tensors()
    want = [run_oracle(kind, cfg, tr, actions[:, n]) for n, tr in enumerate(_traces(env, kind))]
    a_dev = torch.as_tensor(actions, device="cuda:0")
    for t in range(T):
        venv.send_action_tensor(a_dev[t])
        o, r, done = venv.poll_tensors()
        assert o.is_cuda and o.shape == (N, m, env.obs_len) and done == (t == T - 1)
        np.testing.assert_array_equal(o.cpu().numpy(), np.stack([w["obs"][t + 1] for w in want]))
        np.testing.assert_array_equal(r.cpu().numpy(), np.stack([w["reward"][t] for w in want]))


@pytest.mark.parametrize("kind,preset", [("IM", "serial4"), ("IM_div", "div1")])
def test_vector_env_surface(kind, preset):
    cfg = presets.PRESETS[preset]()
    N = 5
    env = ENV_CLASSES[kind](dict(copy_config(cfg), num_envs=N, seed=11))
    venv = BatchedVectorEnv(env)
    assert venv.get_unwrapped() == [env] and venv.num_envs == N
    m, T = env.num_nodes, env.num_periods
    rng = np.random.default_rng(1)
    obs = venv.vector_reset()
    for episode in range(2):
        actions = rng.uniform(-1, 1, size=(T, N, m))
        want = [run_oracle(kind, cfg, tr, actions[:, n]) for n, tr in enumerate(_traces(env, kind))]
        for n in range(N):
            np.testing.assert_array_equal(obs[n], want[n]["obs"][0])
        for t in range(T):
            obs, rew, dones, infos = venv.vector_step([actions[t, n] for n in range(N)])
            assert dones == [t == T - 1] * N and len(infos) == N
            for n in range(N):
                np.testing.assert_array_equal(obs[n], want[n]["obs"][t + 1])
                assert rew[n] == want[n]["reward"][t, 0]
        before = env._episode
        obs = [venv.reset_at(n) for n in range(N)]                # RLlib resets finished envs one by one
        assert env._episode == before + 1                         # ... which is ONE batch reset
    with pytest.raises(ValueError):
        BatchedMultiAgentEnv(env)


def test_step_packed_and_reused_buffers():
    """The lean device-loop entry (step_packed) and the cached per-agent views of reuse_buffers=True: same numbers as the
    allocating path, the returned views alias the reused buffers (overwritten every step, by design)."""
    cfg = presets.serial4()
    N, T, m = 4096, 30, 4
    rng = np.random.default_rng(4)
    demand = rng.poisson(5, size=(N, T)).astype(np.int32)
    actions = torch.as_tensor(rng.uniform(-1, 1, size=(T, N, m)), device="cuda:0")
    fresh = ENV_CLASSES["MAIM"](dict(copy_config(cfg), num_envs=N))
    reuse = ENV_CLASSES["MAIM"](dict(copy_config(cfg), num_envs=N, reuse_buffers=True))
    lean = ENV_CLASSES["MAIM"](dict(copy_config(cfg), num_envs=N, reuse_buffers=True))
    for e in (fresh, reuse, lean):
        e.reset(customer_demand=demand)
    first_views = None
    for t in range(T):
        o1, r1, d1, i1 = fresh.step(actions[t])
        o2, r2, d2, i2 = reuse.step({a: actions[t][:, i] for i, a in enumerate(reuse.agent_names)})      # dict actions too
        o3, r3, d3 = lean.step_packed(actions[t])
        assert d1 == d2 == {"__all__": t == T - 1} and d3 == (t == T - 1) and i1 == i2 == {}
        for i, a in enumerate(fresh.agent_names):
            assert torch.equal(o1[a], o2[a]) and torch.equal(r1[a], r2[a])
            assert torch.equal(o1[a], o3[:, i]) and torch.equal(r1[a], r3[:, i])
        if first_views is None:
            first_views = (o2, r2)
        assert o2 is first_views[0] and r2 is first_views[1]              # cached dicts: nothing rebuilt per step
        assert o2["stage_1"].data_ptr() == reuse.last_obs[:, 1].data_ptr()
