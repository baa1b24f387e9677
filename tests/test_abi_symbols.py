"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every
symbol include/sonics_b200.h declares (no compute calls here -- there is no GPU)."""
import ctypes
import os
import re

import pytest

from sonics_b200 import _abi

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def header_functions():
    text = open(os.path.join(ROOT, "include", "sonics_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sonics_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from sonics_b200 import build
    build.build()
    lib = ctypes.CDLL(_abi.LIB_PATH)
    names = header_functions()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), "missing export: " + n
    assert sorted(_abi.PROTOTYPES) == names, "ctypes prototypes out of sync with the header"


def test_abi_loads_and_reports_version():
    lib = _abi.load()
    assert lib.sonics_abi_version() == 1
    assert lib.sonics_table_bytes(10, 50) > 10 * 48


def test_struct_layouts_match_header():
    assert ctypes.sizeof(_abi.SonicsParams) == 8 * 4 + 10 * 8 + 8 + 8 * 8
    assert ctypes.sizeof(_abi.SonicsVerdict) == 8 * 4 + 9 * 8 + 8 * 8


def test_no_cpu_fallback_without_gpu():
    """On a box without a GPU the engine must fail loudly, not fall back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from sonics_b200.engine import Engine
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Engine(0)
    ctx = ctypes.c_void_p()
    rc = _abi.load().sonics_create(0, ctypes.byref(ctx))
    assert rc == -2 and b"no CPU fallback" in _abi.load().sonics_last_error()


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under sonics_b200/ may reference it."""
    pkg = os.path.join(ROOT, "sonics_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".inl", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower().replace("no oracle", ""), os.path.join(dirpath, f)


def test_tools_and_cli_do_not_use_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may touch oracle/: the measurement tools
    and the CLI must not."""
    import re
    for d in ("tools", "bin"):
        for f in os.listdir(os.path.join(ROOT, d)):
            text = open(os.path.join(ROOT, d, f)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", text, re.M), os.path.join(d, f)
            assert "libsonics_oracle" not in text and "oracle/_ref" not in text, os.path.join(d, f)
