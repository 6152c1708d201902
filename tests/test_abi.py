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
    assert cldrd_lib.cldrd_abi_version() == 2


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


def test_sharded_entry_points_validate_arguments(cldrd_lib):
    """The peer-memory / two-phase entry points reject bad arguments with a code and a message, GPU or not."""
    from cldrd import _lib
    p = C.c_void_p()
    handle = (C.c_ubyte * _lib.PEER_HANDLE_BYTES)()
    assert cldrd_lib.cldrd_peer_alloc(0, 0, C.byref(p), handle) == _lib.E_INVAL
    assert cldrd_lib.cldrd_peer_open(0, None, C.byref(p)) == _lib.E_INVAL
    assert cldrd_lib.cldrd_peer_close(0, C.c_void_p(0x1000)) == _lib.E_INVAL          # not a pointer we handed out
    assert b"cldrd_peer_open" in cldrd_lib.cldrd_last_error()
    assert cldrd_lib.cldrd_peer_close(0, None) == 0 and cldrd_lib.cldrd_peer_free(0, None) == 0
    assert cldrd_lib.cldrd_peer_copy(0, None, None, 16, None) == _lib.E_INVAL
    assert cldrd_lib.cldrd_host_register(None, 0) == _lib.E_INVAL
    assert cldrd_lib.cldrd_host_unregister(None) == 0
    n = C.c_void_p()
    assert cldrd_lib.cldrd_node_create(C.byref(n), 0, 0, 0, 100, 0) == _lib.E_INVAL              # world < 1
    assert cldrd_lib.cldrd_node_create(C.byref(n), 0, 2, 2, 100, 0) == _lib.E_INVAL              # rank outside the world
    assert cldrd_lib.cldrd_node_create(C.byref(n), 0, 2, 0, _lib.MAX_K + 1, 0) == _lib.E_INVAL
    assert cldrd_lib.cldrd_node_create(C.byref(n), 0, 2, 0, 100, 770) == _lib.E_INVAL            # d not a multiple of 4
    assert cldrd_lib.cldrd_node_block_bytes(17, 100, 0) == 0 and cldrd_lib.cldrd_node_block_bytes(8, 1000, 0) > 8 * 1024 * 1000 * 8
    assert cldrd_lib.cldrd_node_block_bytes(8, 1000, 768) - cldrd_lib.cldrd_node_block_bytes(8, 1000, 0) == 8192 * 768 * 4
    assert cldrd_lib.cldrd_node_spread_queries(None, None, 0, 1, None) == _lib.E_INVAL
    assert cldrd_lib.cldrd_node_query_ptr(None, None) == _lib.E_INVAL
    assert cldrd_lib.cldrd_node_set_outputs(None, 1, None, None) == _lib.E_INVAL
    assert cldrd_lib.cldrd_node_search_begin_set(None, None, None, 1, 10, 1, 0, 0, None, None, None) == _lib.E_INVAL
    assert cldrd_lib.cldrd_node_set_wait_mode(None, 0) == _lib.E_INVAL
    assert cldrd_lib.cldrd_host_device_ptr(0, None, None) == _lib.E_INVAL
    assert _lib.MAX_OUT_SETS == 8
    assert cldrd_lib.cldrd_node_search_begin(None, None, None, 1, 10, 1, None, None, None, None, None) == _lib.E_INVAL
    assert cldrd_lib.cldrd_node_search_end(None, None, None, None, 0) == _lib.E_INVAL
    assert cldrd_lib.cldrd_node_attach(None, 0, None, None, -1) == _lib.E_INVAL
    assert cldrd_lib.cldrd_node_detach(None) == 0
    assert cldrd_lib.cldrd_merge_planes(0, None, None, 2, 4, 5, 10, 10, None, None, None, None) == _lib.E_INVAL   # nq > plane_rows
    assert _lib.MAX_PEERS == 16 and _lib.QUERY_BATCH == 8192 and _lib.PEER_HANDLE_BYTES == 72


def test_header_constants_match_python_side():
    import re
    from cldrd import _lib
    text = open(os.path.join(ROOT, "include", "cldrd.h")).read()
    consts = dict(re.findall(r"#define\s+(CLDRD_[A-Z_0-9]+)\s+(-?\d+)", text))
    assert int(consts["CLDRD_MAX_K"]) == _lib.MAX_K
    assert int(consts["CLDRD_SEED_J"]) == _lib.SEED_J
    assert int(consts["CLDRD_MAX_PEERS"]) == _lib.MAX_PEERS
    assert int(consts["CLDRD_QUERY_BATCH"]) == _lib.QUERY_BATCH
    assert int(consts["CLDRD_PEER_HANDLE_BYTES"]) == _lib.PEER_HANDLE_BYTES
    assert int(consts["CLDRD_MAX_OUT_SETS"]) == _lib.MAX_OUT_SETS


def test_integration_md_ctypes_stub_is_live(cldrd_lib, tmp_path):
    """The binding INTEGRATION.md §2 tells a maintainer to add next to retrieval_utils.py is executed as written
    (library path substituted): it must load the library, probe an index file with the documented argument list and
    reach cldrd_shard_create -- which, on this GPU-less box, must refuse loudly instead of falling back to anything."""
    import re
    import numpy as np
    from oracle import flat_ip as O
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    md = open(os.path.join(root, "INTEGRATION.md")).read()
    sec = md[md.index("## 2. The ctypes stub"):md.index("## 3.")]
    code = re.search(r"```python\n(.*?)```", sec, re.S).group(1)
    assert "class B200FlatIP" in code and "/path/to/cl-drd_b200/cldrd/libcldrd.so" in code
    code = code.replace("/path/to/cl-drd_b200/cldrd/libcldrd.so", os.path.join(root, "cl-drd_b200", "cldrd", "libcldrd.so"))
    ns = {}
    exec(compile(code, "INTEGRATION.md#2", "exec"), ns)
    path = str(tmp_path / "t.index")
    O.write_index(path, O.synth(64, 64, 0), O.synth_ids(64))
    import torch
    if torch.cuda.is_available():
        idx = ns["B200FlatIP"](path)
        D, I = idx.search(O.synth(3, 64, 1), 5)
        assert D.shape == (3, 5) and I.shape == (3, 5)
    else:
        ns["B200FlatIP"].__del__ = lambda self: None          # nothing was created
        with pytest.raises(RuntimeError, match="CUDA"):
            ns["B200FlatIP"](path)
    with pytest.raises(RuntimeError, match="cannot open"):
        ns["_ck"](ns["_lib"].cldrd_index_probe(b"/nonexistent.index", None, None, None, None, None, None, None))
