"""CPU-side checks: the C-ABI library loads and exports every symbol include/isla_b200.h declares;
the product never imports the oracle; without a GPU the compute path fails loudly."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from islamcts_b200 import build
    build.build()
    from islamcts_b200 import _lib
    return _lib.load()


def test_every_declared_symbol_is_exported(lib):
    from islamcts_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "isla_b200.h")).read()
    declared = set(re.findall(r"\b(isla_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations found"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/isla_b200.h but not exported"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    assert lib.isla_abi_version() == 3


def test_struct_sizes_match_header(lib):
    import ctypes as C
    from islamcts_b200 import _lib
    assert C.sizeof(_lib.IslaConfig) == 18 * 4 + 7 * 8 + 3 * 8 + 8
    assert C.sizeof(_lib.IslaLayout) == 6 * 4 + 26 * 8


def test_workspace_sizing_and_validation_without_gpu(lib):
    import ctypes as C
    from islamcts_b200 import _lib
    cfg = _lib.IslaConfig(0, 0, 2, 500, 0, 1, 0, 5, 1000, 0, 1, 0, 1000, 0, 0, 0, 0, 0, 1.0, 0.99, 0, 1, 0, 1, 0.7, 0, 0, 65536)
    nbytes = lib.isla_engine_workspace_bytes(C.byref(cfg))
    # 65536 trees x 2004 slots x (32 B record + 16 B state) + path + meta
    assert 6.0e9 < nbytes < 8.0e9
    bad = _lib.IslaConfig(1, 0, 2, 500, 0, 1, 0, 5, 1000, 0, 1, 0, 1000, 0, 0, 0, 0, 0, 1.0, 0.99, 0, 1, 0, 1, 0.7, 0, 0, 16)
    assert lib.isla_engine_workspace_bytes(C.byref(bad)) == _lib.ERR_INVALID   # vanilla on continuous Pendulum
    assert b"discrete" in lib.isla_last_error()
    # CurveEnv: APW on the 2-D box and vanilla on a DiscreteCurveEnv grid size fine; spw/dpw and oversized grids do not
    def curve(agent, n_actions, g0, g1):
        return _lib.IslaConfig(4, agent, n_actions, 50, 0, 1, 0, 5, 1000, 0, 2, 0, 1000, 0, g0, g1, 0, 0, 1.0, 0.99, 0.5, 2, 0, 1, 0.7, 0, 0, 8)
    assert lib.isla_engine_workspace_bytes(C.byref(curve(1, 0, 0, 0))) > 0
    assert lib.isla_engine_workspace_bytes(C.byref(curve(0, 35, 5, 7))) > 0
    assert lib.isla_engine_workspace_bytes(C.byref(curve(0, 34, 5, 7))) == _lib.ERR_INVALID
    assert lib.isla_engine_workspace_bytes(C.byref(curve(0, 0, 0, 0))) == _lib.ERR_INVALID      # vanilla needs the grid
    assert lib.isla_engine_workspace_bytes(C.byref(curve(0, 200, 10, 20))) == _lib.ERR_INVALID
    assert lib.isla_engine_workspace_bytes(C.byref(curve(3, 0, 0, 0))) == _lib.ERR_INVALID       # dpw


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from islamcts_b200 import _lib
    from islamcts_b200.engine import Engine
    assert lib.isla_device_count() == 0
    with pytest.raises(_lib.IslaError):
        Engine(0, n_trees=4)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "islamcts_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{f} imports the oracle"
                assert "isla_oracle" not in src, f"{f} references the oracle library"
