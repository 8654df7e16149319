"""The C ABI from plain C (examples/c_abi_demo.c links libimx_b200.so and the CUDA runtime; no Python, no torch in the
process): one episode of BASELINE config 2's env on replayed traces, per-period imx_step calls and one imx_step_many
call, both checked bit for bit against the C oracle."""
import os
import subprocess

import numpy as np
import pytest

from marl_for_im_b200 import presets
from oracle import c_oracle

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEMO = os.path.join(ROOT, "examples", "c_abi_demo")


@pytest.mark.parametrize("n", [100, 8192])
def test_c_client_matches_oracle(n, tmp_path):
    if not os.path.exists(DEMO):                      # normally prebuilt by __graft_entry__.build(); build it here otherwise
        import __graft_entry__
        __graft_entry__.build()
    assert os.path.exists(DEMO), "examples/c_abi_demo could not be built"
    T, m = 30, 4
    rng = np.random.default_rng(n)
    demand = rng.poisson(5, size=(n, T)).astype(np.int32)
    actions = rng.uniform(-1.1, 1.1, size=(T, n, m))
    fin, fout = tmp_path / "in.bin", tmp_path / "out.bin"
    with open(fin, "wb") as f:
        f.write(np.array([n, T], dtype=np.int64).tobytes())
        f.write(demand.tobytes())
        f.write(actions.tobytes())
    res = subprocess.run([DEMO, str(fin), str(fout)], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stderr + res.stdout
    assert res.stdout.startswith("ok ")
    want = c_oracle.COracle("MAIM", presets.serial4()).run(demand, actions)
    O = want["obs_last"].shape[-1]
    raw = np.fromfile(fout, dtype=np.float64)
    block = T * n * m + n * m * O
    assert raw.size == 2 * block
    for k, how in enumerate(("imx_step x T", "imx_step_many")):
        rew = raw[k * block:k * block + T * n * m].reshape(T, n, m)
        obs = raw[k * block + T * n * m:(k + 1) * block].reshape(n, m, O)
        np.testing.assert_array_equal(rew, want["reward"], err_msg=how)
        np.testing.assert_array_equal(obs, want["obs_last"], err_msg=how)
