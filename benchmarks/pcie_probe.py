"""PCIe probe: pinned D2H / H2D copy-engine bandwidth at the e2e payload sizes of bench.py (context for the e2e number)."""
import json
import time

import torch

dev = torch.device("cuda:0")
out = {}
for name, nbytes in (("d2h_obs_f64_14.7MB", 65536 * 4 * 7 * 8), ("d2h_obs+rew_16.8MB", 65536 * 4 * 8 * 8), ("d2h_obs_f32_7.3MB", 65536 * 4 * 7 * 4),
                     ("d2h_256MB", 256 << 20)):
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    for _ in range(3):
        h.copy_(d, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 20
    for _ in range(reps):
        h.copy_(d, non_blocking=True)
        torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    out[name] = {"ms": dt * 1e3, "GBps": nbytes / dt / 1e9}
for name, nbytes in (("h2d_actions_2.1MB", 65536 * 4 * 8), ("h2d_256MB", 256 << 20)):
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    for _ in range(3):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 20
    for _ in range(reps):
        d.copy_(h, non_blocking=True)
        torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    out[name] = {"ms": dt * 1e3, "GBps": nbytes / dt / 1e9}
print(json.dumps(out))
