"""The C-ABI library loads and exports every symbol include/cldrd.h declares (no compute calls)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "cldrd.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cldrd_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(cldrd_lib):
    from cldrd import _lib
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(cldrd_lib, n), f"{n} declared in cldrd.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, "ctypes table and header disagree"
    assert cldrd_lib.cldrd_abi_version() == 1


def test_no_torch_types_in_header():
    text = open(os.path.join(ROOT, "include", "cldrd.h")).read()
    assert "torch" not in text.lower().replace("pytorch fallback", "")
    assert "at::" not in text and "c10::" not in text


def test_error_reporting_without_gpu(cldrd_lib):
    from cldrd import _lib
    n = C.c_int64()
    rc = cldrd_lib.cldrd_index_probe(b"/nonexistent/x.index", C.byref(n), None, None, None, None, None, None)
    assert rc == _lib.E_IO
    assert b"cannot open" in cldrd_lib.cldrd_last_error()
    rc = cldrd_lib.cldrd_write_run(None, None, None, None, 1, 1, 0, None)
    assert rc == _lib.E_INVAL


def test_compute_entry_points_fail_loudly_without_gpu(cldrd_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from cldrd import _lib
    h = C.c_void_p()
    rc = cldrd_lib.cldrd_shard_create(C.byref(h), 0, 0, 10, 8, _lib.SCAN_SIMT_F32)
    assert rc == _lib.E_CUDA
    assert b"no CPU search path" in cldrd_lib.cldrd_last_error()
    import cldrd
    idx = cldrd.IndexFlatIP(8)
    idx.add(np.zeros((4, 8), dtype=np.float32))
    with pytest.raises(cldrd.CldrdError):
        idx.search(np.zeros((1, 8), dtype=np.float32), 2)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "cl-drd_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src.lower() or "oracle" in f, f"{f} mentions the oracle"
