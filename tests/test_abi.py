"""CPU-only: the C-ABI library loads and exports every symbol include/imx_b200.h declares,
the ctypes struct mirror matches, and the product refuses to run without a GPU."""
import ctypes
import os
import re

import pytest

from harness import ROOT
from marl_for_im_b200 import _lib

HEADER = os.path.join(ROOT, "include", "imx_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(imx_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "run python __graft_entry__.py first"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = declared_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in imx_b200.h but not exported"
    assert set(names) == set(_lib.SYMBOLS), "ctypes binding table and header disagree"


def test_struct_mirror_and_version():
    lib = _lib.load()
    assert lib.imx_config_size() == ctypes.sizeof(_lib.ImxConfig)
    assert lib.imx_abi_version() == _lib.ABI_VERSION == 3


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from marl_for_im_b200 import presets
    from marl_for_im_b200.envs import MultiAgentInvManagement
    with pytest.raises(_lib.ImxError):
        MultiAgentInvManagement(presets.serial4())
    # the raw ABI refuses too
    c = _lib.ImxConfig()
    c.kind, c.num_nodes, c.num_periods, c.prev_length, c.num_envs = 1, 2, 5, 1, 4
    c.a, c.b = -1.0, 1.0
    for i in range(2):
        c.inv_max[i], c.order_max[i], c.delay[i] = 10, 10, 1
    c.price[0], c.price[1], c.price[2] = 3, 2, 1
    h = ctypes.c_void_p()
    rc = _lib.load().imx_create(ctypes.byref(c), ctypes.byref(h))
    assert rc < 0 and b"no CUDA device" in _lib.load().imx_last_error()


def test_config_validation_messages():
    lib = _lib.load()
    c = _lib.ImxConfig()
    c.kind, c.num_nodes, c.num_periods, c.prev_length, c.num_envs = 1, 2, 5, 1, 4
    c.a, c.b = -1.0, 1.0
    for i in range(2):
        c.inv_max[i], c.order_max[i], c.delay[i] = 10, 10, 0          # delay 0 is rejected
    c.price[0], c.price[1], c.price[2] = 3, 2, 1
    h = ctypes.c_void_p()
    assert lib.imx_create(ctypes.byref(c), ctypes.byref(h)) < 0
    assert b"delay" in lib.imx_last_error()
    c.delay[0] = c.delay[1] = 1
    c.time_dependency, c.prev_actions, c.prev_demand = 0, 1, 0           # quirk 3
    assert lib.imx_create(ctypes.byref(c), ctypes.byref(h)) == -5
    assert b"Not Implemented" in lib.imx_last_error()


def test_host_helpers():
    import numpy as np
    from marl_for_im_b200 import topology
    from marl_for_im_b200.spaces import Box
    conn = {0: [1], 1: [2, 3], 2: [4, 5], 3: [], 4: [], 5: []}
    topology.check_connections(conn)
    net = topology.create_network(conn)
    assert topology.get_retailers(net) == [3, 4, 5]
    assert [topology.get_stage(i, net) for i in range(6)] == [0, 1, 2, 2, 3, 3]
    with pytest.raises(Exception):
        topology.check_connections({0: [1], 1: [0]})
    b = Box(low=np.ones(3) * -1, high=np.ones(3), dtype=np.float64, shape=(np.int8(3),))
    assert b.shape == (3,) and isinstance(b.shape[0], int) and b.contains(np.zeros(3))


def _raw_config(kind, cfg, N=65536):
    import numpy as np
    c = _lib.ImxConfig()
    div = kind.endswith("div")
    m = cfg.get("num_nodes", cfg.get("num_stages"))
    c.kind, c.num_nodes, c.num_periods, c.prev_length = _lib.KIND[kind], m, 30, cfg["prev_length"]
    c.time_dependency, c.prev_demand, c.prev_actions = int(cfg["time_dependency"]), int(cfg["prev_demand"]), int(cfg["prev_actions"])
    c.standardise_state, c.standardise_actions = int(cfg.get("standardise_state", True)), int(cfg.get("standardise_actions", True))
    c.independent, c.a, c.b, c.num_envs, c.demand_dist, c.mu = int(cfg.get("independent", False)), -1.0, 1.0, N, 1, 5.0
    for i in range(m):
        c.inv_init[i], c.inv_max[i], c.order_max[i], c.delay[i] = 10, 30, 30, int(cfg["delay"][i])
        c.stock_cost[i], c.backlog_cost[i] = float(cfg["stock_cost"][i]), float(cfg["backlog_cost"][i])
    if div:
        for p, ch in cfg["connections"].items():
            c.num_children[p] = len(ch)
            for k, v in enumerate(ch):
                c.children[p][k] = v
    else:
        for i in range(m + 1):
            c.price[i] = float(cfg["price"][i])
    return c


def test_runtime_specialisation_compiles_for_sm100a_without_gpu():
    """NVRTC builds the specialised kernels from csrc/*.cuh for sm_100a on the CPU box (no launch)."""
    from marl_for_im_b200 import presets
    lib = _lib.load()
    for kind, cfg in (("MAIM", presets.serial4()), ("MAIM", presets.serial8()), ("IM", presets.serial4_dfo()),
                      ("MAIM_div", presets.div1()), ("IM_div", presets.div2(prev_actions=True, prev_length=2))):
        c = _raw_config(kind, cfg)
        for variant in ((0, 1, 2) if kind.startswith("MAIM") else (0, 1)):      # step, no observations, critic rows
            buf = ctypes.create_string_buffer(8192)
            n = lib.imx_jit_compile_check(ctypes.byref(c), variant, buf, 8192)
            assert n > 10000, (kind, variant, n, buf.value.decode()[:500], lib.imx_last_error())


@pytest.mark.parametrize("hints,prefetch", [("0", "0"), ("1", "1"), ("2", "1")])
def test_cache_hint_variants_of_the_specialised_kernels_compile(hints, prefetch, monkeypatch):
    """The L2 eviction-priority modes (IMX_L2_HINTS = none / all / outputs only) and the input prefetch switch select different
    code in the pipelined kernel (createpolicy + cp.async.bulk ... .L2::cache_hint, cp.async.bulk.prefetch.L2): every variant
    must build for sm_100a, serial and divergent (env-per-thread) alike."""
    from marl_for_im_b200 import presets
    monkeypatch.setenv("IMX_L2_HINTS", hints)
    monkeypatch.setenv("IMX_ACT_PREFETCH", prefetch)
    monkeypatch.setenv("IMX_JIT_CACHE", "0")
    lib = _lib.load()
    for kind, cfg in (("MAIM", presets.serial4()), ("MAIM_div", presets.div2())):
        c = _raw_config(kind, cfg)
        for variant in (0, 2):
            buf = ctypes.create_string_buffer(8192)
            n = lib.imx_jit_compile_check(ctypes.byref(c), variant, buf, 8192)
            assert n > 10000, (kind, variant, hints, n, buf.value.decode()[:500], lib.imx_last_error())


def test_header_is_plain_c99_and_links_against_the_library(tmp_path):
    """include/imx_b200.h must be usable from C: compile a C99 translation unit that takes the address of every declared
    entry point with -pedantic -Werror, and link it against libimx_b200.so (no GPU call is made)."""
    import re
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "imx_b200.h")).read()
    names = sorted(set(re.findall(r"\b(imx_[a-z0-9_]+)\s*\(", header)))
    assert len(names) >= 30
    src = tmp_path / "abi.c"
    body = "\n".join(f"    p[{k}] = (void (*)(void))&{n};" for k, n in enumerate(names))
    src.write_text('#include "imx_b200.h"\n#include <stdio.h>\nint main(void) {\n    void (*p[%d])(void);\n%s\n'
                   '    imx_config cfg; (void)cfg;\n    printf("%%d %%d\\n", imx_abi_version(), imx_config_size() == (int)sizeof(imx_config));\n'
                   '    return p[0] == 0;\n}\n' % (len(names), body))
    exe = tmp_path / "abi"
    lib_dir = os.path.join(root, "marl_for_im_b200")
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(root, "include"), str(src), "-o", str(exe),
                    "-L", lib_dir, "-limx_b200", f"-Wl,-rpath,{lib_dir}"], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    assert out == ["3", "1"]                      # ABI version, and the C struct has the size the library was built with
