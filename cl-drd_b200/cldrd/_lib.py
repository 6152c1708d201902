"""ctypes binding of libcldrd.so (include/cldrd.h).  No torch types cross this boundary."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CLDRD_LIB_PATH") or os.path.join(_HERE, "libcldrd.so")   # override: A/B of two builds

SCAN_SIMT_F32, SCAN_TC_TF32, SCAN_TC_F16, SCAN_TC_BF16 = 0, 1, 2, 3
SCAN_NAMES = {"simt": SCAN_SIMT_F32, "tf32": SCAN_TC_TF32, "f16": SCAN_TC_F16, "fp16": SCAN_TC_F16,
              "bf16": SCAN_TC_BF16}
MAX_K = 2048
SEED_J = 32
MAX_PEERS = 16            # CLDRD_MAX_PEERS
QUERY_BATCH = 8192        # CLDRD_QUERY_BATCH
PEER_HANDLE_BYTES = 72    # CLDRD_PEER_HANDLE_BYTES
MAX_OUT_SETS = 8          # CLDRD_MAX_OUT_SETS

E_INVAL, E_IO, E_FORMAT, E_CUDA, E_NOMEM, E_STATE = -1, -2, -3, -4, -5, -6


class CldrdError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libcldrd error {code}: {msg}")
        self.code = code


_lib = None

_c_i64p = C.POINTER(C.c_int64)
_c_i32p = C.POINTER(C.c_int32)
_c_f32p = C.POINTER(C.c_float)

# name -> (restype, argtypes); kept in one table so tests can check it against the header
SIGNATURES = {
    "cldrd_last_error": (C.c_char_p, []),
    "cldrd_abi_version": (C.c_int, []),
    "cldrd_index_probe": (C.c_int, [C.c_char_p, _c_i64p, _c_i32p, _c_i32p, _c_i32p, _c_i32p, _c_i64p, _c_i64p]),
    "cldrd_index_write": (C.c_int, [C.c_char_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32]),
    "cldrd_index_writer_begin": (C.c_int, [C.POINTER(C.c_void_p), C.c_char_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32]),
    "cldrd_index_writer_open_range": (C.c_int, [C.POINTER(C.c_void_p), C.c_char_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int64,
                                                C.c_int64, C.c_int32]),
    "cldrd_index_writer_sync": (C.c_int, [C.c_void_p]),
    "cldrd_index_writer_append": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64]),
    "cldrd_index_writer_finish": (C.c_int, [C.c_void_p, C.c_void_p]),
    "cldrd_index_read_rows": (C.c_int, [C.c_char_p, C.c_int64, C.c_int64, C.c_void_p]),
    "cldrd_index_read_ids": (C.c_int, [C.c_char_p, C.c_int64, C.c_int64, C.c_void_p]),
    "cldrd_shard_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int64, C.c_int64, C.c_int32, C.c_int32]),
    "cldrd_shard_destroy": (None, [C.c_void_p]),
    "cldrd_shard_upload": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64]),
    "cldrd_shard_load_file": (C.c_int, [C.c_void_p, C.c_char_p]),
    "cldrd_shard_adopt": (C.c_int, [C.c_void_p, C.c_void_p]),
    "cldrd_shard_set_ids": (C.c_int, [C.c_void_p, C.c_void_p]),
    "cldrd_shard_finalize": (C.c_int, [C.c_void_p, C.c_void_p]),
    "cldrd_shard_nrows": (C.c_int64, [C.c_void_p]),
    "cldrd_shard_dim": (C.c_int32, [C.c_void_p]),
    "cldrd_shard_scan": (C.c_int32, [C.c_void_p]),
    "cldrd_shard_scan_bytes": (C.c_int64, [C.c_void_p]),
    "cldrd_search_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cldrd_sample_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]),
    "cldrd_seed_from_samples": (C.c_int, [C.c_int, C.c_void_p, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p]),
    "cldrd_search_dev_seeded": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cldrd_verify_seed": (C.c_int, [C.c_int, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cldrd_node_block_bytes": (C.c_int64, [C.c_int32, C.c_int32, C.c_int32]),
    "cldrd_node_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "cldrd_node_query_ptr": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "cldrd_node_spread_queries": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p]),
    "cldrd_node_handle": (C.c_int, [C.c_void_p, C.c_void_p]),
    "cldrd_node_block": (C.c_void_p, [C.c_void_p]),
    "cldrd_node_attach": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32]),
    "cldrd_node_detach": (C.c_int, [C.c_void_p]),
    "cldrd_node_destroy": (None, [C.c_void_p]),
    "cldrd_node_result_ptrs": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "cldrd_node_search_begin": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_void_p]),
    "cldrd_node_set_outputs": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "cldrd_node_search_begin_set": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int64,
                                              C.c_void_p, C.c_void_p, C.c_void_p]),
    "cldrd_node_search_end": (C.c_int, [C.c_void_p, C.c_void_p, _c_i32p, _c_i32p, C.c_int32]),
    "cldrd_node_set_wait_mode": (C.c_int, [C.c_void_p, C.c_int32]),
    "cldrd_node_phase_ms": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    "cldrd_peer_alloc": (C.c_int, [C.c_int, C.c_int64, C.POINTER(C.c_void_p), C.c_void_p]),
    "cldrd_peer_free": (C.c_int, [C.c_int, C.c_void_p]),
    "cldrd_peer_open": (C.c_int, [C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]),
    "cldrd_peer_close": (C.c_int, [C.c_int, C.c_void_p]),
    "cldrd_peer_copy": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "cldrd_shard_norm_bound": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "cldrd_shard_set_norm_bound": (C.c_int, [C.c_void_p, C.c_float]),
    "cldrd_search_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]),
    "cldrd_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_int64]),
    "cldrd_host_register": (C.c_int, [C.c_void_p, C.c_int64]),
    "cldrd_host_unregister": (C.c_int, [C.c_void_p]),
    "cldrd_host_device_ptr": (C.c_int, [C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]),
    "cldrd_host_free": (None, [C.c_void_p]),
    "cldrd_merge_w": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cldrd_merge_planes": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cldrd_merge": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cldrd_shard_last_stats": (C.c_int, [C.c_void_p, _c_i64p]),
    "cldrd_shard_wait_cycles": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64), C.c_int32]),
    "cldrd_shard_set_profiling": (C.c_int, [C.c_void_p, C.c_int32]),
    "cldrd_shard_last_scan_time": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), _c_i64p]),
    "cldrd_shard_last_scan_launches": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), _c_i64p, C.c_int32]),
    "cldrd_scan_dense_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]),
    "cldrd_write_run": (C.c_int, [C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, _c_i64p]),
    "cldrd_write_run_mt": (C.c_int, [C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, _c_i64p]),
    "cldrd_read_run": (C.c_int, [C.c_char_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, _c_i64p, _c_i64p]),
    "cldrd_format_score": (C.c_int, [C.c_float, C.c_char_p]),
    "cldrd_format_score_selfcheck": (C.c_int64, [C.c_uint32, C.c_uint32, C.c_int64, C.POINTER(C.c_uint32), _c_i64p]),
}


def lib():
    """Load libcldrd.so.  There is no fallback: a missing library is a hard error."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CldrdError(E_STATE, f"{LIB_PATH} is missing: build it with `make -C cl-drd_b200` "
                                  f"(or `python -c 'import __graft_entry__ as g; g.build()'`); there is no CPU or "
                                  f"PyTorch fallback for the search path")
    handle = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(handle, name)
        fn.restype = res
        fn.argtypes = args
    _lib = handle
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        msg = lib().cldrd_last_error()
        raise CldrdError(rc, msg.decode("utf-8", "replace") if msg else "unknown error")


def ptr(a) -> C.c_void_p:
    """void* of a numpy array (host) — caller keeps the array alive."""
    return C.c_void_p(a.ctypes.data)


class _PinnedPool:
    """Recycles page-locked result buffers: cudaHostAlloc costs milliseconds per 100 MB, a search
    result is needed for every call.  A buffer returns to the pool when the ndarray that wraps it
    (and every view of it) has been garbage-collected."""

    def __init__(self, max_bytes: int = 4 << 30):
        self.free = {}          # nbytes -> [ptr]
        self.held = 0
        self.max_bytes = max_bytes

    def take(self, nbytes: int) -> int:
        lst = self.free.get(nbytes)
        if lst:
            self.held -= nbytes
            return lst.pop()
        p = C.c_void_p()
        check(lib().cldrd_host_alloc(C.byref(p), nbytes))
        return p.value

    def give(self, ptr_value: int, nbytes: int) -> None:
        if self.held + nbytes > self.max_bytes:
            lib().cldrd_host_free(C.c_void_p(ptr_value))
            return
        self.free.setdefault(nbytes, []).append(ptr_value)
        self.held += nbytes


_pool = _PinnedPool()


class _PinnedOwner:
    def __init__(self, nbytes: int):
        self.nbytes = nbytes
        self.ptr = _pool.take(nbytes)

    def __del__(self):
        try:
            if self.ptr:
                _pool.give(self.ptr, self.nbytes)
                self.ptr = 0
        except Exception:
            pass


def pinned_empty(shape, dtype):
    """numpy array in page-locked memory (falls back to pageable for empty arrays)."""
    import numpy as np
    dtype = np.dtype(dtype)
    n = int(np.prod(shape))
    if n == 0:
        return np.empty(shape, dtype=dtype)
    owner = _PinnedOwner(n * dtype.itemsize)
    buf = (C.c_char * owner.nbytes).from_address(owner.ptr)
    buf._owner = owner                      # the ctypes buffer keeps the pinned block alive
    return np.frombuffer(buf, dtype=dtype, count=n).reshape(shape)
