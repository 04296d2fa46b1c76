"""CPU: the C-ABI library builds, loads and exports every symbol include/mpsim_b200.h declares
(no compute calls: there is no GPU here)."""
import ctypes
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "mpsim_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mpsb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from mpsim_b200.csrc import build
    path = build.build()
    lib = ctypes.CDLL(path)
    names = _header_symbols()
    assert len(names) >= 15
    for name in names:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    from mpsim_b200 import _lib
    assert sorted(_lib.SYMBOLS) == names, "ctypes table and header disagree"
    lib.mpsb_version.restype = ctypes.c_int
    assert lib.mpsb_version() == 100


def test_descriptor_layouts_match_header():
    from mpsim_b200 import _lib
    assert _lib.GATE2_DESC.itemsize == 96
    assert _lib.GATE1_DESC.itemsize == 56
    assert _lib.SITE_REF.itemsize == 24
    assert _lib.GATE2_DESC.fields["bs_site_l"][1] == 48
    assert _lib.GATE1_DESC.fields["chiL"][1] == 48
    assert _lib.SITE_REF.fields["chiL"][1] == 16


def test_argument_errors_without_device():
    """Argument errors are reported before anything is launched (negative return + text)."""
    from mpsim_b200 import _lib
    lib = _lib.load()
    rc = lib.mpsb_apply_gate2(None, 1, 1, 2, 1, 1, 1, 1, 1, None, 0, None, None)
    assert rc < 0
    assert b"descs" in lib.mpsb_last_error()
    assert lib.mpsb_gate2_workspace_bytes(1, 1, 2, 64, 64, 64, 64) >= 128 * 128 * 8


def test_product_path_has_no_cpu_fallback():
    """mpsim_b200 must fail loudly without a CUDA device, never fall back to the oracle."""
    import torch
    import pytest
    import mpsim_b200
    if torch.cuda.is_available():
        pytest.skip("device present")
    with pytest.raises(RuntimeError):
        mpsim_b200.MPS(2)
    src = ""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "mpsim_b200")):
        for f in files:
            if f.endswith(".py"):
                src += open(os.path.join(dirpath, f)).read()
    assert "import oracle" not in src and "from oracle" not in src
    assert "linalg.svd" not in src
