"""Pins oracle/im_oracle.py to the UNMODIFIED reference (build container only; skipped on
the GPU box where /root/reference does not exist — the committed tests/golden fixtures
cover that side)."""
import itertools

import numpy as np
import pytest

from harness import (assert_same, make_delay_mask, random_case, reference_available, run_oracle,
                     run_reference)
from marl_for_im_b200 import presets

pytestmark = pytest.mark.skipif(not reference_available(), reason="reference tree not present")

MODES = list(itertools.product([False, True], repeat=3))   # (td, pd, pa)


def _legal(kind, td, pd, pa):
    return not (kind.startswith("MAIM") and (not td) and pa and (not pd))


@pytest.mark.parametrize("kind,preset", [("MAIM", "serial4"), ("IM", "serial4"), ("MAIM", "serial8"),
                                         ("IM", "serial8"), ("MAIM", "serial2"),
                                         ("MAIM_div", "div1"), ("IM_div", "div1"),
                                         ("MAIM_div", "div2"), ("IM_div", "div2")])
@pytest.mark.parametrize("mode", MODES)
def test_all_obs_modes(kind, preset, mode):
    td, pd, pa = mode
    if not _legal(kind, td, pd, pa):
        pytest.skip("constructor raises 'Not Implemented' (quirk 3)")
    rng = np.random.default_rng(hash((kind, preset, mode)) % (2 ** 32))
    for P, mu, amode, indep in [(1, 5, "uniform", False), (3, 15, "uniform", True), (2, 5, "near_eq", False)]:
        cfg = presets.PRESETS[preset](time_dependency=td, prev_demand=pd, prev_actions=pa, prev_length=P,
                                      independent=indep)
        # heterogeneous capacities / targets so that every per-field maximum is distinguishable
        m = cfg.get("num_nodes", cfg.get("num_stages"))
        cfg["inv_max"] = np.array([30, 25, 40, 35, 30, 45, 20, 30][:m], dtype=float)
        cfg["inv_target"] = np.array([0, 3, 5.5, 1, 0, 2, 4, 0][:m], dtype=float)
        for _ in range(3):
            demand, actions = random_case(kind, cfg, rng, mu=mu, action_mode=amode)
            assert_same(run_reference(kind, cfg, demand, actions), run_oracle(kind, cfg, demand, actions),
                        f"{kind}/{preset}/{mode}/P{P}")


@pytest.mark.parametrize("kind,preset", [("MAIM", "serial4"), ("IM", "serial4"), ("IM_div", "div1")])
def test_non_standardised(kind, preset):
    rng = np.random.default_rng(7)
    for td, pd, pa in MODES:
        if not _legal(kind, td, pd, pa):
            continue
        cfg = presets.PRESETS[preset](time_dependency=td, prev_demand=pd, prev_actions=pa, prev_length=2)
        cfg["standardise_state"] = False
        cfg["standardise_actions"] = False
        for mu in (5, 20):
            demand, actions = random_case(kind, cfg, rng, mu=mu)
            assert_same(run_reference(kind, cfg, demand, actions), run_oracle(kind, cfg, demand, actions),
                        f"{kind} raw {td, pd, pa}")


@pytest.mark.parametrize("kind,preset,std", [("MAIM", "serial4", True), ("MAIM", "serial8", False), ("MAIM_div", "div1", True),
                                             ("MAIM_div", "div2", True), ("IM", "serial4", True), ("IM", "serial4", False),
                                             ("IM_div", "div2", True)])
def test_non_finite_and_huge_actions(kind, preset, std):
    """+-inf, NaN (MAIM kinds) and finite values either side of 2^63: the MAIM kinds convert to int64 BEFORE clipping
    (MAIM_env.py:344-347), so on the reference's x86-64 everything outside [-2^63, 2^63) becomes INT64_MIN and clips to order 0;
    the IM kinds clip first (IM_env.py:300-302).  The oracle follows the reference through every special value, live."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    from make_golden import inject_nonfinite
    rng = np.random.default_rng(hash((kind, preset, std)) % (2 ** 32))
    for _ in range(4):
        cfg = presets.PRESETS[preset]()
        if not std:
            cfg["standardise_actions"] = False
        demand, actions = random_case(kind, cfg, rng)
        actions = inject_nonfinite(actions, kind, rng)
        with np.errstate(all="ignore"):
            want = run_reference(kind, cfg, demand, actions, None)
        got = run_oracle(kind, cfg, demand, actions, None)
        assert_same(want, got)


@pytest.mark.parametrize("kind,preset", [("MAIM", "serial4"), ("IM", "serial8"), ("MAIM_div", "div2"),
                                         ("IM_div", "div1")])
def test_noisy_delay_replayed(kind, preset):
    rng = np.random.default_rng(11)
    cfg = presets.PRESETS[preset]()
    for _ in range(6):
        demand, actions = random_case(kind, cfg, rng, mu=6, action_mode="near_eq")
        mask = make_delay_mask(kind, cfg["delay"], cfg["num_periods"], 0.3, rng)
        assert_same(run_reference(kind, cfg, demand, actions, mask), run_oracle(kind, cfg, demand, actions, mask),
                    f"{kind} noisy delay")


def test_custom_ab_and_order_max():
    rng = np.random.default_rng(3)
    cfg = presets.serial4(prev_actions=True)
    cfg["a"], cfg["b"] = 0, 3
    cfg["order_max"] = np.array([20., 30., 25., 30.])
    for kind in ("MAIM", "IM"):
        T, m = 30, 4
        demand = rng.poisson(7, T)
        actions = rng.uniform(-0.2, 3.3, size=(T, m))
        assert_same(run_reference(kind, cfg, demand, actions), run_oracle(kind, cfg, demand, actions), kind)


def test_base_stock_rollout_and_dfo():
    from scipy.stats import poisson
    from oracle import im_oracle
    from oracle.ref_import import load_reference
    R = load_reference()
    rng = np.random.default_rng(5)
    for preset, zs in (("serial4", [25, 25, 25, 25]), ("serial8", [12.5, 20, 31.25, 8, 25, 25, 17, 40])):
        cfg = presets.PRESETS[preset](time_dependency=False, prev_demand=False, prev_actions=False,
                                      standardise_state=False, standardise_actions=False)
        ref_env = R.InvManagement(dict(cfg))
        ref_env.reset()                      # sets env.dist / dist_param (poisson, mu=5)
        orc = im_oracle.OracleEnv("IM", cfg)
        for _ in range(4):
            demand = rng.poisson(5, 30)
            z = np.array(zs, dtype=float)
            want = R.dfo_func(z, ref_env, demand)
            pmf = poisson.pmf(demand, mu=5)
            got = im_oracle.dfo_value(orc, z, demand, pmf)
            assert got == want
            rewards = im_oracle.base_stock_rollout(orc, z, demand)
            ref_env.reset(customer_demand=demand)
            for t in range(30):
                _, r, _, _ = ref_env.step(R.base_stock_policy(z, ref_env))
                assert r == rewards[t]
            np.testing.assert_allclose([im_oracle.poisson_pmf(int(k), 5.0) for k in demand], pmf, rtol=1e-13)


@pytest.mark.parametrize("kind", ["MAIM_div", "IM_div"])
def test_random_divergent_networks(kind):
    """Random trees (5-12 nodes, up to 5 children per node): the round-robin split with more than two children."""
    from harness import random_tree_config
    rng = np.random.default_rng(2027 if kind == "MAIM_div" else 2028)
    for trial in range(12):
        m = int(rng.integers(5, 13))
        cfg = random_tree_config(rng, m, int(rng.integers(2, 6)), periods=16, prev_actions=bool(trial % 2), prev_length=1 + trial % 3,
                                 independent=bool(trial % 3 == 0), share_network=(kind == "MAIM_div" and trial % 4 == 0))
        for amode in ("near_eq", "uniform"):
            demand, actions = random_case(kind, cfg, rng, mu=4, action_mode=amode)
            assert_same(run_reference(kind, cfg, demand, actions), run_oracle(kind, cfg, demand, actions), f"{kind} tree {cfg['connections']}")


def test_named_configurations_match_hyperparams_py():
    """presets.named_env_config reproduces the env_config of every named configuration in hyperparams.py."""
    import sys
    from oracle.ref_import import load_reference
    load_reference()                                    # puts the reference tree on sys.path
    import hyperparams
    for name in presets.NAMED_CONFIGS:
        want = hyperparams.get_hyperparams(name)
        cls, got = presets.named_env_config(name)
        assert {"InventoryManagement": "InvManagement", "MultiAgentInventoryManagement": "MultiAgentInvManagement"}[want["env"]] == cls
        assert set(got) == set(want["env_config"]), name
        for k, v in want["env_config"].items():
            np.testing.assert_array_equal(np.asarray(got[k]), np.asarray(v), err_msg=f"{name}.{k}")
    with pytest.raises(KeyError):
        presets.named_env_config("MA_13")


def test_topology_helpers_match_utils_py():
    """marl_for_im_b200.topology against utils.py:87-130 on random trees."""
    from harness import random_tree_config
    from marl_for_im_b200 import topology
    from oracle.ref_import import load_reference
    R = load_reference()
    rng = np.random.default_rng(9)
    for _ in range(25):
        m = int(rng.integers(3, 14))
        conn = random_tree_config(rng, m, int(rng.integers(1, 6)))["connections"]
        R.check_connections(conn)
        topology.check_connections(conn)
        want = R.create_network(conn)
        got = topology.create_network(conn)
        np.testing.assert_array_equal(got, want)
        assert topology.get_retailers(got) == R.get_retailers(want)
        assert [topology.get_stage(i, got) for i in range(m)] == [R.get_stage(i, want) for i in range(m)]
    bad = {0: [1], 2: [1], 1: []}
    with pytest.raises(Exception):
        R.check_connections(bad)
    with pytest.raises(Exception):
        topology.check_connections(bad)


@pytest.mark.parametrize("kind", ["MAIM", "IM"])
def test_random_serial_chains(kind):
    """Random chains (2-12 stages, lead times up to 5, histories up to 4, rescale intervals other than [-1, 1])."""
    from harness import random_serial_config
    rng = np.random.default_rng(4040 if kind == "MAIM" else 4041)
    done = 0
    while done < 20:
        cfg = random_serial_config(rng, int(rng.integers(2, 13)))
        if not _legal(kind, cfg["time_dependency"], cfg["prev_demand"], cfg["prev_actions"]):
            continue
        T, m = cfg["num_periods"], cfg["num_stages"]
        demand = rng.poisson(rng.uniform(3, 12), T)
        if cfg["standardise_actions"]:
            span = cfg["b"] - cfg["a"]
            actions = rng.uniform(cfg["a"] - 0.1 * span, cfg["b"] + 0.1 * span, size=(T, m))
        else:
            actions = rng.uniform(-3, 50, size=(T, m))
        assert_same(run_reference(kind, cfg, demand, actions), run_oracle(kind, cfg, demand, actions), f"{kind} {cfg}")
        done += 1


def test_dfo_func_divergent_and_noisy_live():
    """The reference's own call (inv_management_div.py:228: dfo_func on InvManagementDiv with demand [R, T]) and dfo_func after a
    noisy reset (the flag is sticky, MAIM_env.py:192-194), live against the oracle on random trees and chains."""
    from scipy.stats import poisson
    from harness import _uniform_replayer, copy_config, random_tree_config
    from oracle import im_oracle
    from oracle.ref_import import load_reference
    import warnings
    R = load_reference()
    rng = np.random.default_rng(77)
    for trial in range(10):
        m = int(rng.integers(3, 9))
        cfg = random_tree_config(rng, m, int(rng.integers(1, 4)), periods=int(rng.integers(5, 41)), time_dependency=False, prev_demand=False)
        cfg.update(standardise_state=False, standardise_actions=False, demand_dist="poisson", mu=5, delay=np.asarray(cfg["delay"], dtype=np.int64))
        noisy = bool(trial % 2)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            env = R.InvManagementDiv(copy_config(cfg))
            nret, T = len(env.retailers), env.num_periods
            orc = im_oracle.OracleEnv("IM_div", copy_config(cfg))
            for _ in range(3):
                z = rng.integers(5, 40, size=m).astype(float) + rng.choice([0.0, 0.25], size=m)
                demand = rng.poisson(5, size=(nret, T))
                mask = rng.uniform(size=(T, m)) <= 0.3 if noisy else None
                saved = np.random.uniform
                try:
                    if noisy:
                        env.reset(customer_demand=demand, noisy_delay=True, noisy_delay_threshold=0.5)
                        seq = iter(_uniform_replayer("IM_div", [int(d) for d in env.delay], T, mask))
                        np.random.uniform = lambda *a, **k: next(seq)
                    want = R.dfo_func(z, env, demand)
                finally:
                    np.random.uniform = saved
                assert im_oracle.dfo_value(orc, z, demand, poisson.pmf(demand, mu=5), mask) == want, (trial, cfg["connections"])
