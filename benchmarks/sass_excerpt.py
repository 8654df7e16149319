#!/usr/bin/env python
"""profiles/r2_sass_step_kernel_pipe_excerpt.txt: SASS evidence of the NVRTC-specialised pipelined step kernel, produced on the
build box (no GPU): resource usage, opcode histogram and the TMA / mbarrier / prefetch / cache-hint instructions.

    python benchmarks/sass_excerpt.py > profiles/r2_sass_step_kernel_pipe_excerpt.txt
"""
import ctypes
import os
import re
import subprocess
import sys
import tempfile
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from marl_for_im_b200 import _lib, presets  # noqa: E402
from marl_for_im_b200.envs import ENV_CLASSES  # noqa: E402

FUN = "_ZN3imx16step_kernel_pipeILi4ELi3ELi1ELi1ELb0EEEvNS_8StepArgsENS_10TileLayoutENS_8PipeArgsE"
KEEP = re.compile(r"UBLKCP|UBLKPF|SYNCS|UTMACMDFLUSH|ACQBULK|DEPBAR|PREEXIT|FENCE\.VIEW|NANOSLEEP|CCTL|UTMAPF")


def dump(n_envs):
    lib = _lib.load()
    c = ENV_CLASSES["MAIM"](dict(presets.serial4(), num_envs=n_envs, _config_only=True)).imx_config
    path = tempfile.mktemp(suffix=".cubin")
    os.environ["IMX_JIT_DUMP"] = path
    os.environ["IMX_JIT_CACHE"] = "0"
    buf = ctypes.create_string_buffer(4096)
    assert lib.imx_jit_compile_check(ctypes.byref(c), 0, buf, 4096) > 0, buf.value
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", FUN, path], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", path], capture_output=True, text=True).stdout
    os.unlink(path)
    return sass, res


def main():
    for n_envs, what in ((65536, "65536 envs: output-only L2 priorities"), (262144, "262144 envs: full L2 eviction priorities")):
        sass, res = dump(n_envs)
        ins = [l for l in sass.splitlines() if re.search(r"/\*[0-9a-f]{4}\*/", l)]
        print(f"# SASS of the NVRTC-specialised persistent pipelined step kernel (config 2: MAIM 4-stage, MA_6 obs mode, {what})")
        print(f"# IMX_JIT_DUMP=<file> imx_jit_compile_check(cfg, 0, ...)  then  cuobjdump -sass -fun {FUN}")
        print("# (a) resource usage")
        lines = res.splitlines()
        for i, l in enumerate(lines):
            if "step_kernel_pipe" in l or "step_kernel_tmaI" in l:
                print(" " + l.strip())
                print("  " + lines[i + 1].strip())
        ops = Counter()
        for l in ins:
            m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", l)
            if m:
                ops[m.group(1)] += 1
        print(f"# (b) opcode histogram of the kernel ({sum(ops.values())} instructions)")
        for k, v in ops.most_common(24):
            print(f"    {v:4d} {k}")
        print("# (c) the TMA / mbarrier instructions (UBLKCP = cp.async.bulk [.L2::cache_hint when a fourth operand is present], UBLKPF = "
              "cp.async.bulk.prefetch.L2, SYNCS = mbarrier ops, UTMACMDFLUSH = bulk commit, ACQBULK = griddepcontrol.wait, PREEXIT = "
              "griddepcontrol.launch_dependents, DEPBAR = wait_group)")
        for i, l in enumerate(ins):
            if KEEP.search(l):
                print(f"{i}: {l.rstrip()[:150]}")
        print()


if __name__ == "__main__":
    main()
