#!/usr/bin/env python
"""N-rank PCIe probe (run under torchrun, one rank per GPU): what can this box move between the GPUs and host memory when
ALL ranks transfer at once?  Per rank and in aggregate, with NUMA-local pinned buffers (dist.bind_host_to_gpu):
  d2h_ce     device -> pinned host through the copy engine (cudaMemcpyAsync), the e2e payload of one step (obs + reward)
  d2h_sm     the same bytes written by SM stores into the mapped pinned buffer (what the zero-copy step kernel does)
  h2d_ce     pinned host -> device, the actions of one step
  duplex     d2h_ce and h2d_ce on two streams at once
One JSON line (rank 0).  profiles/r2_pcie_probe_*rank.json are the committed runs.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 benchmarks/pcie_probe_nrank.py
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from marl_for_im_b200.dist import bind_host_to_gpu  # noqa: E402
from marl_for_im_b200.envs import _DevView  # noqa: E402


def main():
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    cores = bind_host_to_gpu(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    d2h_bytes = 65536 * 4 * 8 * 8                     # obs [N,4,7] + reward [N,4] float64 of config 2: 16.8 MB
    h2d_bytes = 65536 * 4 * 8                         # actions [N,4] float64: 2.1 MB
    d_out = torch.empty(d2h_bytes, dtype=torch.uint8, device=dev).fill_(1)
    h_out = torch.empty(d2h_bytes, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(h2d_bytes, dtype=torch.uint8, device=dev)
    h_in = torch.empty(h2d_bytes, dtype=torch.uint8).pin_memory().fill_(2)
    # device-side alias of the pinned output buffer (UVA: same address): an SM copy kernel then stores over PCIe
    h_out_alias = torch.as_tensor(_DevView(h_out.data_ptr(), (d2h_bytes // 16, 2), "<f8"), device=dev)
    d_out_f = d_out.view(torch.float64).reshape(-1, 2)
    s2 = torch.cuda.Stream()
    reps = 30

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
            torch.cuda.synchronize()                  # one transfer per step, like imx_step_host
        dt = (time.perf_counter() - t0) / reps
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def duplex():
        with torch.cuda.stream(s2):
            d_in.copy_(h_in, non_blocking=True)
        h_out.copy_(d_out, non_blocking=True)

    res = {}
    for name, fn, nbytes in (("d2h_ce", lambda: h_out.copy_(d_out, non_blocking=True), d2h_bytes),
                             ("d2h_sm", lambda: h_out_alias.copy_(d_out_f), d2h_bytes),
                             ("h2d_ce", lambda: d_in.copy_(h_in, non_blocking=True), h2d_bytes),
                             ("duplex", duplex, d2h_bytes + h2d_bytes)):
        dt = timed(fn)
        res[name] = {"ms_max_over_ranks": dt * 1e3, "GBps_per_rank": nbytes / dt / 1e9, "GBps_aggregate": world * nbytes / dt / 1e9, "bytes": nbytes}
    if rank == 0:
        try:
            import subprocess
            topo = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
            numa = sorted({ln.split()[-2] for ln in topo.splitlines() if ln.startswith("GPU") and len(ln.split()) > 3})
        except Exception:
            numa = None
        print(json.dumps({"world": world, "host_cores": os.cpu_count(), "cores_bound_rank0": cores, "numa_affinity_column": numa, "probe": res}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
