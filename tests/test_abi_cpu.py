"""CPU tests of the drop-in boundary: libdlmcq.so builds, loads, and exports exactly the entry
points include/dlmcq.h declares (no compute calls - there is no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dlmcq.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dlmcq_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def libpath():
    from dlmc_quant_b200 import build
    return build.build()


def test_header_declares_entry_points():
    names = declared_symbols()
    assert "dlmcq_fq_forward" in names and "dlmcq_fq_backward" in names and len(names) >= 25


def test_library_exports_every_declared_symbol(libpath):
    h = ctypes.CDLL(libpath)
    missing = [n for n in declared_symbols() if not hasattr(h, n)]
    assert not missing, f"declared in dlmcq.h but not exported: {missing}"


def test_ctypes_table_matches_header(libpath):
    from dlmc_quant_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    h = _lib.lib()
    assert h.dlmcq_version() == 100
    assert h.dlmcq_status_string(0) == b"ok"
    assert h.dlmcq_status_string(-3) == b"workspace too small"


def test_workspace_query_needs_no_gpu(libpath):
    from dlmc_quant_b200 import _lib
    h = _lib.lib()
    base = h.dlmcq_workspace_bytes(None)
    assert base >= 256 + 2048 * 80 * 4          # 80 sweep partials per CTA fit
    lay = _lib.Layout(64, 2048, 49, 0)
    assert h.dlmcq_workspace_bytes(ctypes.byref(lay)) >= 256 + 64 * 2048 * 16
    ctx = ctypes.c_void_p()
    assert h.dlmcq_host_ctx_create(ctypes.byref(ctx), 12345) == -1          # chunk must be a multiple of 8
    assert h.dlmcq_host_ctx_synchronize(None) == -1 and h.dlmcq_host_ctx_destroy(None) == 0


def test_argument_validation_without_gpu(libpath):
    """Bad arguments are rejected before any CUDA call is made."""
    from dlmc_quant_b200 import _lib
    h = _lib.lib()
    lay = _lib.Layout(1, 1, 16, 0)
    qp = _lib.QParams(1, 0, 15, 0.0, None, None)
    assert h.dlmcq_fq_forward(None, None, None, ctypes.byref(lay), ctypes.byref(qp), None) == -1
    bad = _lib.Layout(1, 0, 16, 0)
    assert h.dlmcq_fq_forward(None, None, None, ctypes.byref(bad), ctypes.byref(qp), None) == -1


def test_product_path_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from dlmc_quant_b200 import functional as F
    from dlmc_quant_b200._lib import DlmcqError
    with pytest.raises(DlmcqError):
        F.fq_forward(torch.zeros(8), torch.ones(1), None, 0, 15, 1)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under dlmc_quant_b200/ may reference it."""
    pkg = os.path.join(ROOT, "dlmc_quant_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("no oracle", ""), f"{f} mentions the oracle"
