"""The Python oracle against the committed reference-generated fixtures (runs anywhere)."""
import numpy as np
import pytest

from harness import assert_same, golden_names, load_golden, run_oracle
from oracle import im_oracle


@pytest.mark.parametrize("name", golden_names())
def test_oracle_matches_golden(name):
    g = load_golden(name)
    got = run_oracle(g["kind"], g["config"], g["demand_trace"], g["actions"], g["delay_mask"])
    assert_same(g["ref"], got, name)


def test_golden_set_is_complete():
    assert len(golden_names()) >= 24


def test_np_sum_order_matches_numpy():
    rng = np.random.default_rng(1)
    for n in range(1, 34):
        for _ in range(50):
            a = rng.uniform(-5, 5, n) * rng.uniform(0, 100, n)
            assert im_oracle.np_sum_order(list(a)) == np.sum(a)


def test_quirk3_not_implemented():
    from marl_for_im_b200 import presets
    cfg = presets.serial4(time_dependency=False, prev_demand=False, prev_actions=True)
    with pytest.raises(Exception, match="Not Implemented"):
        im_oracle.OracleEnv("MAIM", cfg)
    im_oracle.OracleEnv("IM", cfg)       # the single-agent env supports this mode


def test_topology_helpers():
    parent, children, depth, retailers = im_oracle.topology({0: [1], 1: [2, 3], 2: [4, 5], 3: [], 4: [], 5: []}, 6)
    assert parent == [-1, 0, 1, 1, 2, 2]
    assert depth == [0, 1, 2, 2, 3, 3]
    assert retailers == [3, 4, 5]
    with pytest.raises(ValueError):
        im_oracle.topology({0: [1], 1: [0]}, 2)
