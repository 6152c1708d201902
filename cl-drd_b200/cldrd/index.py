"""The subset of the `faiss` Python API that CL-DRD's retriever scripts touch, backed by
libcldrd.so.  Same names, argument meaning and error behaviour:

    IndexFlatIP(d), IndexIDMap(index), IndexIDMap2(index), index_factory(d, "Flat", METRIC_INNER_PRODUCT)
        retriever/index_text.py:91-97, retriever/retrieval_utils.py:119-128
    write_index / read_index
        retriever/index_text.py:105, retriever/retrieve_top_passages.py:85
    StandardGpuResources, GpuClonerOptions, index_cpu_to_gpu,
    GpuResourcesVector, IntVector, GpuMultipleClonerOptions, index_cpu_to_gpu_multiple
        retriever/retrieval_utils.py:155-184
    index.search(x, k) -> (D float32 [n,k], I int64 [n,k])
        retriever/retrieval_utils.py:135,143

There is no CPU search: a host-side index that is asked to search places itself on the current
CUDA device first.  Without the CUDA library / a B200 every search raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from typing import List, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import CldrdError, check, lib, ptr

METRIC_INNER_PRODUCT = 0
METRIC_L2 = 1


def default_scan() -> str:
    """Scan mode used when the caller does not pick one: CLDRD_SCAN env var, else "auto"
    (fp16 scan when the values fit fp16, else tf32).  Every mode returns exact fp32 results."""
    return os.environ.get("CLDRD_SCAN", "auto").lower()


# ---------------------------------------------------------------------------------------------
# host-side index objects
# ---------------------------------------------------------------------------------------------


class _RowStore:
    """fp32 rows either in host memory (blocks appended by add) or still in an index file."""

    def __init__(self, d: int):
        self.d = int(d)
        self.blocks: List[np.ndarray] = []
        self.n = 0
        self.file: Optional[str] = None  # lazily loaded file-backed rows
        self.file_n = 0

    def add(self, x: np.ndarray) -> None:
        if self.file is not None:
            self.materialize()
        self.blocks.append(np.ascontiguousarray(x, dtype=np.float32))
        self.n += x.shape[0]

    def materialize(self) -> np.ndarray:
        if self.file is not None:
            out = np.empty((self.file_n, self.d), dtype=np.float32)
            check(lib().cldrd_index_read_rows(self.file.encode(), 0, self.file_n, ptr(out)))
            self.blocks = [out]
            self.file = None
        if len(self.blocks) != 1:
            self.blocks = [np.concatenate(self.blocks, axis=0) if self.blocks else
                           np.empty((0, self.d), dtype=np.float32)]
        return self.blocks[0]


def _is_cuda_tensor(x) -> bool:
    return (not isinstance(x, np.ndarray)) and getattr(x, "is_cuda", False) is True


def _check_xd(x, d: int, what: str):
    """Device-resident queries (extension): a float32 [n, d] CUDA tensor."""
    import torch
    if x.dtype != torch.float32:
        raise TypeError(f"{what}: tensor must be float32, got {x.dtype}")
    assert x.dim() == 2, f"{what}: expected a 2-D tensor"
    assert x.shape[1] == d, f"{what}: dimension {x.shape[1]} != index dimension {d}"
    return x.contiguous()


def _check_x(x, d: int, what: str) -> np.ndarray:
    # faiss' SWIG wrapper asserts on shape and raises TypeError on dtype; same here
    if not isinstance(x, np.ndarray):
        raise TypeError(f"{what}: expected a numpy.ndarray, got {type(x).__name__}")
    if x.dtype != np.float32:
        raise TypeError(f"{what}: array must be float32, got {x.dtype}")
    assert x.ndim == 2, f"{what}: expected a 2-D array"
    assert x.shape[1] == d, f"{what}: dimension {x.shape[1]} != index dimension {d}"
    return np.ascontiguousarray(x)


class Index:
    """Common behaviour of the host-side flat / id-map objects."""

    d: int
    metric_type = METRIC_INNER_PRODUCT
    is_trained = True

    def __init__(self):
        self._gpu: Optional["GpuIndexFlat"] = None

    def _invalidate(self):
        if self._gpu is not None:
            self._gpu.close()
            self._gpu = None

    def search(self, x, k):
        """No CPU search exists: clone to the current CUDA device on first use, then search there."""
        if self._gpu is None:
            import torch  # device plumbing only
            if not torch.cuda.is_available():
                raise CldrdError(_lib.E_CUDA, "no CUDA device: this index has no CPU search path")
            self._gpu = index_cpu_to_gpu(StandardGpuResources(), torch.cuda.current_device(), self)
        return self._gpu.search(x, k)


class IndexFlatIP(Index):
    def __init__(self, d: int):
        super().__init__()
        self.d = int(d)
        self._rows = _RowStore(d)

    @property
    def ntotal(self) -> int:
        return self._rows.n

    def add(self, x):
        x = _check_x(x, self.d, "add")
        self._invalidate()
        self._rows.add(x)

    def reset(self):
        self._invalidate()
        self._rows = _RowStore(self.d)

    # helpers for the cloner / writer
    def _ids(self) -> Optional[np.ndarray]:
        return None


class IndexIDMap(Index):
    _fourcc2 = False

    def __init__(self, index: IndexFlatIP):
        super().__init__()
        if not isinstance(index, IndexFlatIP):
            raise TypeError("IndexIDMap wraps an IndexFlatIP here (the only case the retriever uses)")
        assert index.ntotal == 0, "index must be empty on input"
        self.index = index
        self.d = index.d
        self._id_blocks: List[np.ndarray] = []
        self._ids_file: Optional[str] = None

    @property
    def ntotal(self) -> int:
        return self.index.ntotal

    @property
    def _rows(self) -> _RowStore:
        return self.index._rows

    def add_with_ids(self, x, ids):
        x = _check_x(x, self.d, "add_with_ids")
        if not isinstance(ids, np.ndarray):
            raise TypeError("add_with_ids: ids must be a numpy.ndarray")
        if ids.dtype != np.int64:
            raise TypeError(f"add_with_ids: ids must be int64, got {ids.dtype}")
        assert ids.shape == (x.shape[0],), "add_with_ids: not same nb of vectors as ids"
        self._invalidate()
        self._load_ids()
        self.index._rows.add(x)
        self._id_blocks.append(np.ascontiguousarray(ids))

    def add(self, x):
        raise RuntimeError("add not implemented for this type of index (use add_with_ids)")

    def _load_ids(self):
        if self._ids_file is not None:
            n = self.index._rows.file_n if self.index._rows.file is not None else self.index.ntotal
            out = np.empty((n,), dtype=np.int64)
            check(lib().cldrd_index_read_ids(self._ids_file.encode(), 0, n, ptr(out)))
            self._id_blocks = [out]
            self._ids_file = None

    def _ids(self) -> np.ndarray:
        self._load_ids()
        if len(self._id_blocks) != 1:
            self._id_blocks = [np.concatenate(self._id_blocks) if self._id_blocks else
                               np.empty((0,), dtype=np.int64)]
        return self._id_blocks[0]

    @property
    def id_map(self) -> np.ndarray:
        return self._ids()


class IndexIDMap2(IndexIDMap):
    _fourcc2 = True


def index_factory(d: int, description: str, metric: int = METRIC_L2):
    """Only what retriever/retrieval_utils.py:119 asks for: a flat inner-product index."""
    if description != "Flat":
        raise RuntimeError(f"index_factory: only 'Flat' is supported here, got {description!r}")
    if metric != METRIC_INNER_PRODUCT:
        raise RuntimeError("index_factory: only METRIC_INNER_PRODUCT is supported here")
    return IndexFlatIP(d)


# ---------------------------------------------------------------------------------------------
# index files
# ---------------------------------------------------------------------------------------------


def write_index(index, path: str) -> None:
    if isinstance(index, (GpuIndexFlat, GpuIndexShards)):
        raise RuntimeError("write_index: clone the index to CPU first (faiss behaves the same)")
    rows = index._rows.materialize()
    ids = index._ids()
    check(lib().cldrd_index_write(str(path).encode(), ptr(rows), ptr(ids) if ids is not None else None,
                                  rows.shape[0], index.d, 1 if getattr(index, "_fourcc2", False) else 0))


def read_index(path: str):
    """Parse the headers only; rows stay in the file until they are cloned to a GPU (straight
    from the file into HBM) or until a host-side operation needs them."""
    path = str(path)
    n, d, metric = C.c_int64(), C.c_int32(), C.c_int32()
    has_ids, idmap2 = C.c_int32(), C.c_int32()
    doff, ioff = C.c_int64(), C.c_int64()
    check(lib().cldrd_index_probe(path.encode(), C.byref(n), C.byref(d), C.byref(metric), C.byref(has_ids),
                                  C.byref(idmap2), C.byref(doff), C.byref(ioff)))
    flat = IndexFlatIP(d.value)
    wrapped = (IndexIDMap2 if idmap2.value else IndexIDMap)(flat) if has_ids.value else None
    flat._rows.file = path
    flat._rows.file_n = n.value
    flat._rows.n = n.value
    if wrapped is None:
        return flat
    wrapped._ids_file = path
    return wrapped


# ---------------------------------------------------------------------------------------------
# GPU side
# ---------------------------------------------------------------------------------------------


class StandardGpuResources:
    """Kept for signature compatibility; libcldrd sizes its own workspace (about 1 GiB per shard)."""

    def __init__(self):
        self.temp_memory = None

    def setTempMemory(self, nbytes: int):
        self.temp_memory = int(nbytes)

    def noTempMemory(self):
        self.temp_memory = 0


class GpuClonerOptions:
    def __init__(self):
        self.useFloat16 = False  # True: fp16 scan copy.  Results are exact fp32 either way.
        self.scan = None         # extension: "simt" | "tf32" | "f16" | "bf16" | "auto"


class GpuMultipleClonerOptions(GpuClonerOptions):
    def __init__(self):
        super().__init__()
        self.shard = False


class _Vector(list):
    def push_back(self, v):
        self.append(v)

    def size(self):
        return len(self)

    def at(self, i):
        return self[i]


class GpuResourcesVector(_Vector):
    pass


class IntVector(_Vector):
    pass


def _resolve_scan(co) -> str:
    scan = getattr(co, "scan", None) if co is not None else None
    if scan is None:
        scan = "f16" if (co is not None and getattr(co, "useFloat16", False)) else default_scan()
    return scan


class _Shard:
    """Owns one cldrd_shard handle."""

    def __init__(self, device: int, row0: int, nrows: int, d: int, scan: str):
        self.handle = C.c_void_p()
        self.device, self.row0, self.nrows, self.d = int(device), int(row0), int(nrows), int(d)
        self.scan_request = scan
        self._create("f16" if scan == "auto" else scan)
        self._keepalive = None

    def _create(self, scan: str):
        if scan not in _lib.SCAN_NAMES:
            raise ValueError(f"unknown scan mode {scan!r}")
        check(lib().cldrd_shard_create(C.byref(self.handle), self.device, self.row0, self.nrows, self.d,
                                       _lib.SCAN_NAMES[scan]))

    def close(self):
        if self.handle:
            lib().cldrd_shard_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def fill(self, filler) -> None:
        """filler(shard) populates rows/ids; finalize, retrying with tf32 when "auto" picked fp16
        for values that do not fit fp16."""
        self._filler = filler
        filler(self)
        try:
            check(lib().cldrd_shard_finalize(self.handle, None))
        except CldrdError as e:
            if self.scan_request != "auto" or e.code != _lib.E_INVAL:
                raise
            self.refill("tf32")

    def refill(self, scan: str) -> None:
        """Rebuild the shard with another scan precision (rows are loaded / adopted again)."""
        self.close()
        self._create(scan)
        self._filler(self)
        check(lib().cldrd_shard_finalize(self.handle, None))

    def stats(self) -> dict:
        arr = (C.c_int64 * 8)()
        check(lib().cldrd_shard_last_stats(self.handle, arr))
        names = ["launches", "chunks", "fallback_queries", "rescored", "survivors", "max_list", "tc_tiles",
                 "exact_compactions"]
        return dict(zip(names, list(arr)))

    def set_profiling(self, on: bool) -> None:
        check(lib().cldrd_shard_set_profiling(self.handle, 1 if on else 0))

    def scan_time(self):
        """(summed device ms of the scan kernels of the last search, number of scan launches)."""
        ms, n = C.c_double(), C.c_int64()
        check(lib().cldrd_shard_last_scan_time(self.handle, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def wait_cycles(self, reset: bool = True) -> dict:
        """MMA-issuer wait breakdown of the profiled filter scans (fractions of the issuer's lifetime)."""
        arr = (C.c_uint64 * 4)()
        check(lib().cldrd_shard_wait_cycles(self.handle, arr, 1 if reset else 0))
        tot = max(int(arr[3]), 1)
        return {"operands": arr[0] / tot, "tmem_stage": arr[1] / tot, "unit_id": arr[2] / tot, "cycles": int(arr[3])}

    def scan_launches(self):
        """[(rows, ms)] of every scan launch of the last search (profiling on)."""
        ms = (C.c_double * 256)()
        rows = (C.c_int64 * 256)()
        n = lib().cldrd_shard_last_scan_launches(self.handle, ms, rows, 256)
        return [(int(rows[j]), float(ms[j])) for j in range(max(n, 0))]

    @property
    def scan(self) -> str:
        code = lib().cldrd_shard_scan(self.handle)
        return {0: "simt", 1: "tf32", 2: "f16", 3: "bf16"}[code]

    @property
    def scan_bytes(self) -> int:
        return int(lib().cldrd_shard_scan_bytes(self.handle))


def _fill_from_index(index, row0: int, nrows: int, with_ids: bool):
    """Returns a filler that moves rows [row0,row0+nrows) of a host-side index into a shard."""
    store = index._rows

    def filler(sh: _Shard):
        if store.file is not None:
            check(lib().cldrd_shard_load_file(sh.handle, store.file.encode()))  # loads ids too
            if not with_ids:
                check(lib().cldrd_shard_set_ids(sh.handle, None))
            return
        rows = store.materialize()
        part = rows[row0:row0 + nrows]
        check(lib().cldrd_shard_upload(sh.handle, ptr(part), 0, nrows))
        ids = index._ids() if with_ids else None
        if ids is not None:
            part_ids = np.ascontiguousarray(ids[row0:row0 + nrows])
            check(lib().cldrd_shard_set_ids(sh.handle, ptr(part_ids)))

    return filler


class GpuIndexFlat:
    """What index_cpu_to_gpu returns: the whole index resident on one GPU."""

    def __init__(self, shard: _Shard, ntotal: int, d: int):
        self._shard = shard
        self.ntotal = ntotal
        self.d = d
        self.metric_type = METRIC_INNER_PRODUCT
        self.is_trained = True
        self._lock = threading.Lock()

    def close(self):
        self._shard.close()

    def search(self, x, k):
        """x: float32 [n, d] numpy array (the reference's call) or, as an extension, a CUDA tensor holding the
        embeddings the encoder just produced (no host round trip on the way in).  Returns numpy (D, I)."""
        k = int(k)
        on_dev = _is_cuda_tensor(x)
        x = _check_xd(x, self.d, "search") if on_dev else _check_x(x, self.d, "search")
        assert k > 0, "search: k must be positive"
        if k > _lib.MAX_K:
            raise RuntimeError(f"search: k={k} above the GPU limit {_lib.MAX_K} (same limit as faiss GpuIndexFlat)")
        n = x.shape[0]
        # results land in page-locked arrays the caller owns: one DMA, no staging copy
        D = _lib.pinned_empty((n, k), np.float32)
        I = _lib.pinned_empty((n, k), np.int64)
        if n and on_dev:
            import torch
            Dd, Id = self.search_device(x, k)
            st = torch.cuda.current_stream(x.device)
            check(lib().cldrd_peer_copy(x.device.index, ptr(D), C.c_void_p(Dd.data_ptr()), D.nbytes, C.c_void_p(st.cuda_stream)))
            check(lib().cldrd_peer_copy(x.device.index, ptr(I), C.c_void_p(Id.data_ptr()), I.nbytes, C.c_void_p(st.cuda_stream)))
            st.synchronize()
        elif n:
            with self._lock:
                check(lib().cldrd_search_host(self._shard.handle, ptr(x), n, k, ptr(D), ptr(I)))
        return D, I

    def search_device(self, q, k: int, translate_ids: bool = True):
        """Extension: torch CUDA tensor in, torch CUDA tensors out (no host copies)."""
        import torch
        assert q.is_cuda and q.dtype == torch.float32 and q.dim() == 2 and q.shape[1] == self.d
        q = q.contiguous()
        n = q.shape[0]
        D = torch.empty((n, k), dtype=torch.float32, device=q.device)
        I = torch.empty((n, k), dtype=torch.int64, device=q.device)
        if n:
            stream = torch.cuda.current_stream(q.device).cuda_stream
            with self._lock:
                check(lib().cldrd_search_dev(self._shard.handle, C.c_void_p(q.data_ptr()), n, int(k),
                                             1 if translate_ids else 0, C.c_void_p(D.data_ptr()),
                                             C.c_void_p(I.data_ptr()), C.c_void_p(stream)))
        return D, I

    def sample_device(self, q, k: int):
        """Step 1 of the sharded protocol: best SEED_J sample scan scores per query, [n, SEED_J]."""
        import torch
        q = q.contiguous()
        n = q.shape[0]
        out = torch.empty((n, _lib.SEED_J), dtype=torch.float32, device=q.device)
        stream = torch.cuda.current_stream(q.device).cuda_stream
        with self._lock:
            check(lib().cldrd_sample_dev(self._shard.handle, C.c_void_p(q.data_ptr()), n, int(k),
                                         C.c_void_p(out.data_ptr()), C.c_void_p(stream)))
        return out

    def search_device_seeded(self, q, k: int, seed=None, translate_ids: bool = False):
        """Step 3: search with an external seed threshold per query (None = unseeded progressive).
        Returns (D, I, eps2) with eps2 = 2*eps per query for the caller's verification."""
        import torch
        q = q.contiguous()
        n = q.shape[0]
        D = torch.empty((n, k), dtype=torch.float32, device=q.device)
        I = torch.empty((n, k), dtype=torch.int64, device=q.device)
        eps2 = torch.empty((n,), dtype=torch.float32, device=q.device)
        if n:
            stream = torch.cuda.current_stream(q.device).cuda_stream
            with self._lock:
                check(lib().cldrd_search_dev_seeded(
                    self._shard.handle, C.c_void_p(q.data_ptr()), n, int(k), 1 if translate_ids else 0,
                    C.c_void_p(seed.data_ptr()) if seed is not None else None, C.c_void_p(D.data_ptr()),
                    C.c_void_p(I.data_ptr()), C.c_void_p(eps2.data_ptr()), C.c_void_p(stream)))
        return D, I, eps2

    def last_stats(self) -> dict:
        return self._shard.stats()

    @property
    def scan(self) -> str:
        return self._shard.scan


def index_cpu_to_gpu(res, device: int, index, co: Optional[GpuClonerOptions] = None) -> GpuIndexFlat:
    """retriever/retrieval_utils.py:163.  File-backed indexes stream straight into HBM."""
    n, d = index.ntotal, index.d
    sh = _Shard(device, 0, n, d, _resolve_scan(co))
    sh.fill(_fill_from_index(index, 0, n, with_ids=isinstance(index, IndexIDMap)))
    return GpuIndexFlat(sh, n, d)


def shard_ranges(ntotal: int, parts: int) -> List[range]:
    """Contiguous passage-row ranges, rows_r = [floor(r*N/G), floor((r+1)*N/G))  (SURVEY §8e)."""
    return [range((r * ntotal) // parts, ((r + 1) * ntotal) // parts) for r in range(parts)]


class GpuIndexShards:
    """What index_cpu_to_gpu_multiple(..., shard=True) returns inside ONE process: the index row-sharded over
    several GPUs.  A search is the node-wide protocol of include/cldrd.h ("Sharded search on one node") driven by
    one host thread: one asynchronous cldrd_node_search_begin per shard and batch, the shards' kernels exchanging
    sample scores, counts and re-scored lists through each other's HBM over NVLink, every shard's merge kernel
    storing its slice of the result straight into the caller's page-locked arrays.  Results are bit-identical to
    the single-GPU search."""

    SEED_MIN_ROWS = 1 << 20
    RING = 3

    def __init__(self, shards: List[_Shard], ids: Optional[np.ndarray], ntotal: int, d: int):
        import torch
        self._torch = torch
        self._shards = shards
        self.ntotal, self.d = ntotal, d
        self.metric_type = METRIC_INNER_PRODUCT
        self.is_trained = True
        self._lock = threading.Lock()
        self._devs = [torch.device("cuda", sh.device) for sh in shards]
        self._streams = [torch.cuda.Stream(device=dv) for dv in self._devs]
        self._nodes: List[C.c_void_p] = []
        self._node_k = 0
        self._q = [None] * len(shards)
        self.last_seed_misses = 0
        # every decision of the protocol uses the same error band on all shards: same scan precision, same norm bound
        if len({sh.scan for sh in shards}) > 1:
            for sh in shards:
                if sh.scan != "tf32":
                    sh.refill("tf32")
        bound = 0.0
        for sh in shards:
            b = C.c_float()
            check(lib().cldrd_shard_norm_bound(sh.handle, C.byref(b)))
            bound = max(bound, b.value)
        for sh in shards:
            check(lib().cldrd_shard_set_norm_bound(sh.handle, C.c_float(bound)))
        self._id_maps = [None] * len(shards)
        if ids is not None:
            first = {}
            for i, dv in enumerate(self._devs):      # one replica per device
                if dv not in first:
                    first[dv] = torch.from_numpy(ids).to(dv)
                self._id_maps[i] = first[dv]

    @property
    def scan(self) -> str:
        return self._shards[0].scan

    def last_stats(self) -> List[dict]:
        return [sh.stats() for sh in self._shards]

    def _drop_nodes(self):
        for n in self._nodes:
            lib().cldrd_node_detach(n)
        for n in self._nodes:
            lib().cldrd_node_destroy(n)
        self._nodes = []
        self._node_k = 0

    def _ensure_nodes(self, k: int):
        if self._nodes and k <= self._node_k:
            return
        self._drop_nodes()
        G = len(self._shards)
        try:
            for r, sh in enumerate(self._shards):
                h = C.c_void_p()
                check(lib().cldrd_node_create(C.byref(h), sh.device, G, r, int(k), self.d if self.d % 4 == 0 else 0))
                self._nodes.append(h)
            for r, sh in enumerate(self._shards):
                for p, other in enumerate(self._shards):
                    if p != r:
                        blk = lib().cldrd_node_block(self._nodes[p])
                        check(lib().cldrd_node_attach(self._nodes[r], p, None, C.c_void_p(blk), other.device))
        except Exception:
            self._drop_nodes()
            raise
        self._node_k = int(k)

    def close(self):
        self._drop_nodes()
        for s in self._shards:
            s.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _end_all(self, b0: int):
        again = None
        for sh, node in zip(self._shards, self._nodes):
            nfail = C.c_int32()
            idx = (C.c_int32 * _lib.QUERY_BATCH)()
            check(lib().cldrd_node_search_end(sh.handle, node, C.byref(nfail), idx, _lib.QUERY_BATCH))
            mine = [b0 + idx[i] for i in range(nfail.value)]
            assert again is None or again == mine, "the shards disagree on the queries to search again"
            again = mine
        return again

    def _run(self, qs, n: int, k: int, seeded: bool, outs, out_rows=None):
        """qs[i]: the queries on shard i's device (a tensor, or a callable (b0, nb) -> device address that may enqueue
        whatever brings the batch there); outs[i]: (scores, ids) base addresses as shard i's device sees the output
        arrays; out_rows[i]: optional int32 device tensor with the output row of every query."""
        inflight, again = [], []
        for b0 in range(0, n, _lib.QUERY_BATCH):
            nb = min(_lib.QUERY_BATCH, n - b0)
            if len(inflight) >= self.RING:
                again += self._end_all(inflight.pop(0))
            qptr = [qs[i](b0, nb) if callable(qs[i]) else qs[i][b0:b0 + nb].data_ptr() for i in range(len(self._shards))]
            for i, (sh, node) in enumerate(zip(self._shards, self._nodes)):
                oD, oI = outs[i]
                if out_rows is None:
                    oD, oI, rows = oD + b0 * k * 4, oI + b0 * k * 8, None
                else:
                    rows = C.c_void_p(out_rows[i][b0:b0 + nb].data_ptr())
                idm = C.c_void_p(self._id_maps[i].data_ptr()) if self._id_maps[i] is not None else None
                check(lib().cldrd_node_search_begin(sh.handle, node, C.c_void_p(qptr[i]), nb, k,
                                                    1 if seeded else 0, C.c_void_p(oD), C.c_void_p(oI), rows, idm,
                                                    C.c_void_p(self._streams[i].cuda_stream)))
            inflight.append(b0)
        while inflight:
            again += self._end_all(inflight.pop(0))
        return again

    def search(self, x, k):
        torch = self._torch
        k = int(k)
        on_dev = _is_cuda_tensor(x)
        x = _check_xd(x, self.d, "search") if on_dev else _check_x(x, self.d, "search")
        assert k > 0, "search: k must be positive"
        if k > _lib.MAX_K:
            raise RuntimeError(f"search: k={k} above the GPU limit {_lib.MAX_K}")
        n = x.shape[0]
        # results land in page-locked arrays the caller owns, written by the shards' merge kernels
        D = _lib.pinned_empty((n, k), np.float32)
        I = _lib.pinned_empty((n, k), np.int64)
        if n == 0:
            return D, I
        with self._lock:
            self._ensure_nodes(k)
            G = len(self._shards)
            spread = (not on_dev) and self.d % 4 == 0 and G > 1
            qs, outs = [], []
            if spread:
                # host queries: every device uploads 1/G of the batch over its own PCIe link into its exchange block and
                # stores it into the other blocks over NVLink (cldrd_node_spread_queries) -- not G uploads of the same bytes
                row_bytes, src = self.d * 4, x.ctypes.data

                def make_src(i):
                    sh, node, dv = self._shards[i], self._nodes[i], self._devs[i]
                    qx = C.c_void_p()
                    check(lib().cldrd_node_query_ptr(node, C.byref(qx)))
                    st = C.c_void_p(self._streams[i].cuda_stream)

                    def q_src(b0, nb):
                        sl = (nb + G - 1) // G
                        lo = min(nb, i * sl)
                        hi = min(nb, lo + sl)
                        if hi > lo:
                            check(lib().cldrd_peer_copy(dv.index, C.c_void_p(qx.value + lo * row_bytes),
                                                        C.c_void_p(src + (b0 + lo) * row_bytes), (hi - lo) * row_bytes, st))
                        check(lib().cldrd_node_spread_queries(sh.handle, node, lo, hi - lo, st))
                        return qx.value
                    return q_src

                qs = [make_src(i) for i in range(G)]
                q0 = None
            else:
                # one upload (none for device-resident embeddings), then NVLink copies to the other shards
                q0 = x.to(self._devs[0]) if on_dev else torch.from_numpy(x).to(self._devs[0])
                torch.cuda.current_stream(q0.device).synchronize()
            for i, dv in enumerate(self._devs):
                if not spread:
                    with torch.cuda.stream(self._streams[i]):
                        if dv == self._devs[0]:
                            self._streams[i].wait_stream(torch.cuda.current_stream(dv))
                            qi = q0
                        else:
                            qi = torch.empty_like(q0, device=dv)
                            qi.copy_(q0, non_blocking=True)
                        qs.append(qi)
                pD, pI = C.c_void_p(), C.c_void_p()
                check(lib().cldrd_host_device_ptr(dv.index, C.c_void_p(D.ctypes.data), C.byref(pD)))
                check(lib().cldrd_host_device_ptr(dv.index, C.c_void_p(I.ctypes.data), C.byref(pI)))
                outs.append((pD.value, pI.value))
            again = self._run(qs, n, k, self.ntotal >= self.SEED_MIN_ROWS, outs)
            self.last_seed_misses = len(again)
            if again:   # rare: seed above the true k-th score, or a survivor buffer overflowed: unseeded retry
                q2, rows = [], []
                for i, dv in enumerate(self._devs):
                    with torch.cuda.stream(self._streams[i]):
                        idx = torch.tensor(again, dtype=torch.int64, device=dv)
                        q2.append(torch.from_numpy(np.ascontiguousarray(x[again])).to(dv) if spread else qs[i][idx].contiguous())
                        rows.append(idx.to(torch.int32))
                left = self._run(q2, len(again), k, False, outs, rows)
                assert not left, "an unseeded batch cannot raise queries"
            for st in self._streams:
                st.synchronize()
        return D, I


def index_cpu_to_gpu_multiple(vres, vdev: Sequence[int], index, co: Optional[GpuMultipleClonerOptions] = None):
    """retriever/retrieval_utils.py:182.  shard=True: rows split over the devices; shard=False
    (replicas) would only waste HBM bandwidth for this workload, so it is served by the same
    sharded layout."""
    devs = list(vdev)
    assert len(devs) >= 1
    n, d = index.ntotal, index.d
    scan = _resolve_scan(co)
    shards = []
    for dev, rr in zip(devs, shard_ranges(n, len(devs))):
        sh = _Shard(dev, rr.start, len(rr), d, scan)
        sh.fill(_fill_from_index(index, rr.start, len(rr), with_ids=False))
        shards.append(sh)
    ids = index._ids() if isinstance(index, IndexIDMap) else None
    return GpuIndexShards(shards, ids, n, d)


def index_gpu_to_cpu(index):
    raise RuntimeError("index_gpu_to_cpu: keep the host-side index you cloned from")
