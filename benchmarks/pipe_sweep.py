#!/usr/bin/env python
"""Plain one-tile-per-CTA step kernel vs. the persistent pipelined one (IMX_PIPE) over batch size, ring depth, resident CTAs
per SM and CTA size.  One JSON line per point."""
import argparse
import itertools
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "benchmarks"))

from bench_configs import time_steps  # noqa: E402
from floor_sweep import CONFIGS  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="serial4")
    ap.add_argument("--envs", default="16384,32768,65536,131072,262144")
    ap.add_argument("--threads", default="128")
    ap.add_argument("--stages", default="3,4")
    ap.add_argument("--ctas", default="2,3,4")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--no-plain", action="store_true")
    args = ap.parse_args()
    for name in args.configs.split(","):
        kind, preset = CONFIGS[name]
        for thr in args.threads.split(","):
            os.environ["IMX_TMA_THREADS"] = thr
            for n in (int(x) for x in args.envs.split(",")):
                points = [] if args.no_plain else [("0", "", "")]
                points += [("1", s, c) for s, c in itertools.product(args.stages.split(","), args.ctas.split(","))]
                for pipe, st, ct in points:
                    os.environ["IMX_PIPE"] = pipe
                    if st:
                        os.environ["IMX_PIPE_STAGES"] = st
                        os.environ["IMX_PIPE_CTAS"] = ct
                    try:
                        r = time_steps(kind, preset(), n, args.reps)
                    except Exception as exc:
                        r = {"error": str(exc)[:300]}
                    r.update(config=name, tma_threads=int(thr), envs=n, pipe=pipe, stages=st, ctas=ct)
                    print(json.dumps(r), flush=True)


if __name__ == "__main__":
    main()
