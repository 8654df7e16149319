"""Loader of tests/golden/dfo/dfo_cases.npz (reference-generated dfo_func values, see make_golden_dfo.py)."""
import json
import os

import numpy as np

from harness import GOLDEN_DIR

PATH = os.path.join(GOLDEN_DIR, "dfo", "dfo_cases.npz")


def _cfg(raw):
    cfg = {}
    for k, v in json.loads(raw).items():
        if isinstance(v, dict) and "__nd__" in v:
            cfg[k] = np.array(v["__nd__"], dtype=v["dtype"])
        elif isinstance(v, dict) and "__dict__" in v:
            cfg[k] = {int(a): list(b) for a, b in v["__dict__"].items()}
        else:
            cfg[k] = v
    return cfg


def load_dfo_cases():
    z = np.load(PATH, allow_pickle=False)
    out = []
    for name in [str(n) for n in z["names"]]:
        meta = json.loads(str(z[name + "__meta"]))
        out.append(dict(name=name, kind=meta["kind"], noisy=meta["noisy"], config=_cfg(meta["config"]), z=z[name + "__z"],
                        demand=z[name + "__demand"], mask=z[name + "__mask"], dfo=z[name + "__dfo"]))
    return out
