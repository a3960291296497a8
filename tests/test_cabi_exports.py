"""CPU: libia_b200.so loads without a GPU and exports exactly the entry points include/ia_b200.h declares; the
product fails loudly (no CPU fallback) when handed CPU tensors."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "ia_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ia_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    from instantavatar_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "run `python -c 'import __graft_entry__ as g; g.build()'` first"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/ia_b200.h but not exported"
    assert sorted(_lib.SYMBOLS) == syms, (sorted(set(syms) - set(_lib.SYMBOLS)), sorted(set(_lib.SYMBOLS) - set(syms)))
    assert lib.ia_abi_version() == 1


def test_host_only_entry_points_work_without_gpu():
    from instantavatar_b200 import _lib
    lay = _lib.hashgrid_layout()
    assert lay["total"] == 6513496 and lay["res"][0] == 16 and lay["res"][3] == 54 and lay["res"][-1] == 7007
    assert lay["size"][4] == 1 << 19 and lay["offset"][1] == 4096
    # invalid arguments are reported through the error channel, not by crashing
    rc = _lib.lib().ia_set_option(b"no_such_option", ctypes.c_int(1))
    assert rc == -1 and b"unknown option" in _lib.lib().ia_last_error()


def test_no_cpu_fallback():
    import torch
    from instantavatar_b200 import ops
    x = torch.zeros((4, 3))
    with pytest.raises(RuntimeError, match="CUDA tensors"):
        ops.ngp_forward(ops.Scene(table_h=x, mlp_h=x, net_center=x[0], net_scale=x[0]), x)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "instantavatar_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith(".py"):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f"{f} imports oracle/"


def test_every_entry_point_is_documented_for_integrators():
    """INTEGRATION.md must name every function include/ia_b200.h declares (the binding guide is part of the boundary)"""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "ia_b200.h")).read()
    doc = open(os.path.join(root, "INTEGRATION.md")).read()
    syms = sorted(set(re.findall(r"\b(ia_[a-z0-9_]+)\s*\(", hdr)))
    assert len(syms) >= 50
    missing = [s for s in syms if s not in doc]
    assert not missing, missing
