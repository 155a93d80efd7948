"""CPU-side checks of the drop-in boundary: libb2r.so loads without a GPU, exports every symbol
include/b2r.h declares, and the ctypes table covers exactly those symbols; argument validation of
the entry points that need no device work."""
import ctypes as C
import os
import re

import pytest

from msra_practice_project_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "b2r.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b2r_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported_and_bound():
    syms = header_symbols()
    assert len(syms) >= 15
    lib = _lib.lib()
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/b2r.h but not exported by libb2r.so"
    assert sorted(_lib.SIGNATURES) == syms, "ctypes table and header disagree"


def test_version_and_size_queries():
    lib = _lib.lib()
    assert lib.b2r_version() == 100
    assert lib.b2r_mlp_tc_packed_bytes(0) > 1187840
    assert lib.b2r_mlp_f32_workspace_bytes(0, 10, 1) == 10 * 2516 * 4
    assert lib.b2r_mlp_f32_workspace_bytes(1, 10, 1) == 10 * 4616 * 4
    assert lib.b2r_mlp_f32_workspace_bytes(0, 10 ** 7, 0) == 65536 * 2516 * 4      # inference is chunked
    assert lib.b2r_mlp_f32_bwd_scratch_bytes(0, 3) == 3 * 512 * 4


def test_bad_arguments_return_negative_and_set_message():
    lib = _lib.lib()
    rc = lib.b2r_composite_fwd(None, None, None, 3, 1, 1, None, None, None, None, None)
    assert rc < 0 and b"NULL" in lib.b2r_last_error()
    rc = lib.b2r_sample_pdf(None, 0, None, 0, None, 1, 4, 4, None, 0, None, None, None, None)
    assert rc < 0
    inp = _lib.MlpInput()
    rc = lib.b2r_mlp_tc_fwd(0, 16, 1, C.byref(inp), 16, 0, None)
    assert rc < 0 and b"exactly one" in lib.b2r_last_error()
    with pytest.raises(RuntimeError):
        _lib.check(rc, "b2r_mlp_tc_fwd")
