"""The C-ABI library loads and exports every symbol include/robir_b200.h declares (no compute without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    from robir_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return _lib


def test_header_symbols_exported(built):
    hdr = open(os.path.join(ROOT, "include", "robir_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(robir_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 20
    handle = ctypes.CDLL(built.LIB_PATH)
    for name in declared:
        assert hasattr(handle, name), "missing export: " + name
    assert sorted(built.EXPORTED) == declared, "ctypes binding and header disagree"
    assert handle.robir_abi_version() == 1


def test_struct_layouts_match_header(built):
    # ctypes mirrors of the POD argument blocks: sizes as the C compiler lays them out (LP64)
    assert ctypes.sizeof(built.SgParams) == 4 * 4 + 35 * 8
    assert ctypes.sizeof(built.SdfParams) == 8 + 4 + 3 * 4 + 8 * 8 * 2 + 7 * 8
    assert ctypes.sizeof(built.OctreeView) == 2 * 8 + 4 * 4 + 6 * 4
    assert ctypes.sizeof(built.OctCastParams) == ctypes.sizeof(built.OctreeView) + 3 * 8 + 3 * 4 + 3 * 4 + 6 * 8


def test_product_fails_loudly_without_cuda(built):
    import torch
    from robir_b200 import ops
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(built.RobirError):
        ops.pe_linear(torch.zeros(4, 3), torch.zeros(64, 256), None)
