"""Row-sharded search with one process per GPU (torchrun), the B200-native form of the
reference's `index_cpu_to_gpu_multiple(..., shard=True)` (retriever/retrieval_utils.py:174-182).

rank r keeps passage rows shard_ranges(N, G)[r] resident in its HBM; queries are replicated.
Per query batch (include/cldrd.h "Sharded search on one node") every rank makes ONE asynchronous
call: sample scan -> sample scores stored into every rank's block -> flag barrier -> levels and the global filter
seed -> fused scan + filter -> candidates counted against the levels, counts stored into every rank's block ->
barrier -> only what can still reach the global top-k is re-scored in fp32 and stored straight into the HBM of
the rank that merges the query (peer memory over NVLink, mapped with CUDA IPC) -> barrier -> every rank merges
and verifies its slice (same u64 key order as the single-GPU search: bit-identical results) and stores it where
the result is wanted: rank 0's HBM or a shared page-locked host block.  No NCCL call and no host synchronisation
inside a batch.  Without peer mapping NCCL moves the lists (all-gather, all-to-all, gather).  torch.distributed is
plumbing only; no collective touches the index.
"""
from __future__ import annotations

import ctypes as C
import os
import time
from typing import Optional, Tuple

import torch
import torch.distributed as dist

from . import _lib
from ._lib import CldrdError, check, lib
from .index import GpuIndexFlat, _Shard, shard_ranges


def gather_candidates(D_local: torch.Tensor, I_local: torch.Tensor, dst: int = 0, group=None
                      ) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
    """Gather per-rank [nq,k] candidate lists to `dst` as [world, nq, k].  Works for CUDA tensors
    over NCCL and for CPU tensors over gloo (used by the CPU tests of this plumbing)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if world == 1:
        return D_local.unsqueeze(0), I_local.unsqueeze(0)
    if rank == dst:
        allD = torch.empty((world,) + tuple(D_local.shape), dtype=D_local.dtype, device=D_local.device)
        allI = torch.empty((world,) + tuple(I_local.shape), dtype=I_local.dtype, device=I_local.device)
        dist.gather(D_local.contiguous(), list(allD.unbind(0)), dst=dst, group=group)
        dist.gather(I_local.contiguous(), list(allI.unbind(0)), dst=dst, group=group)
        return allD, allI
    dist.gather(D_local.contiguous(), None, dst=dst, group=group)
    dist.gather(I_local.contiguous(), None, dst=dst, group=group)
    return None, None


def merge_candidates(allD: torch.Tensor, allI: torch.Tensor, id_map: Optional[torch.Tensor] = None,
                     k: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """[P,nq,w] scores + global rows (CUDA, -1 padded) -> merged [nq,k] via the libcldrd merge kernel
    (k defaults to w)."""
    assert allD.is_cuda and allD.dim() == 3 and allI.shape == allD.shape
    P, n, w = allD.shape
    k = w if k is None else int(k)
    allD, allI = allD.contiguous(), allI.contiguous()
    outD = torch.empty((n, k), dtype=torch.float32, device=allD.device)
    outI = torch.empty((n, k), dtype=torch.int64, device=allD.device)
    st = torch.cuda.current_stream(allD.device).cuda_stream
    check(lib().cldrd_merge_w(allD.device.index, C.c_void_p(allD.data_ptr()), C.c_void_p(allI.data_ptr()), P, n, w, k,
                              C.c_void_p(id_map.data_ptr()) if id_map is not None else None,
                              C.c_void_p(outD.data_ptr()), C.c_void_p(outI.data_ptr()), C.c_void_p(st)))
    return outD, outI


def _coll_dev(group) -> torch.device:
    """Where tensors of the (setup-time) collectives live: the GPU under NCCL, host memory under gloo (the
    CPU tests, and two ranks sharing one GPU, which NCCL refuses)."""
    if dist.get_backend(group) == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def _all_ok(mine: bool, group) -> bool:
    flag = torch.tensor([1 if mine else 0], dtype=torch.int32, device=_coll_dev(group))
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    return bool(flag.item())


class _NodeExchange:
    """This rank's cldrd_node: its exchange block plus the blocks of the node's other ranks mapped into this
    process (CUDA IPC handles travel through the process group once, at construction).  Everything a search
    exchanges afterwards is stored by kernels into these blocks (include/cldrd.h, "Sharded search on one node")."""

    def __init__(self, device: int, rank: int, world: int, max_k: int, group, d: int = 0):
        self.device, self.rank, self.world, self.group = device, rank, world, group
        self.max_k = max_k
        self.d = d if d % 4 == 0 else 0        # query buffer for host-side queries (cldrd_node_spread_queries)
        self.handle = C.c_void_p()
        ok = True
        mine = torch.zeros((_lib.PEER_HANDLE_BYTES,), dtype=torch.uint8)
        try:
            check(lib().cldrd_node_create(C.byref(self.handle), device, world, rank, max_k, self.d))
            buf = (C.c_ubyte * _lib.PEER_HANDLE_BYTES)()
            check(lib().cldrd_node_handle(self.handle, buf))
            mine = torch.frombuffer(bytearray(buf), dtype=torch.uint8)
        except Exception:
            ok = False
        cdev = _coll_dev(group)
        allh = torch.empty((world * _lib.PEER_HANDLE_BYTES,), dtype=torch.uint8, device=cdev)   # flat: gloo insists
        dist.all_gather_into_tensor(allh, mine.to(cdev), group=group)
        allh = allh.cpu().view(world, _lib.PEER_HANDLE_BYTES)
        ok = _all_ok(ok, group)           # nobody maps a block whose owner failed to create it
        if ok:
            try:
                for r in range(world):
                    if r != rank:
                        hb = (C.c_ubyte * _lib.PEER_HANDLE_BYTES).from_buffer_copy(bytes(allh[r].tolist()))
                        check(lib().cldrd_node_attach(self.handle, r, hb, None, -1))
            except Exception:
                ok = False
            ok = _all_ok(ok, group)
        self.ok = ok
        if not ok:
            self.close()

    def query_ptr(self) -> int:
        p = C.c_void_p()
        check(lib().cldrd_node_query_ptr(self.handle, C.byref(p)))
        return p.value

    def result_ptrs(self, owner: int):
        d, i = C.c_void_p(), C.c_void_p()
        check(lib().cldrd_node_result_ptrs(self.handle, owner, C.byref(d), C.byref(i)))
        return d.value, i.value

    def phase_ms(self) -> dict:
        arr = (C.c_double * 7)()
        check(lib().cldrd_node_phase_ms(self.handle, arr))
        return dict(zip(["sample+barrier+levels", "scan+select", "counts+barrier+rescore", "barrier+merge+store",
                         "barrier+status", "idle_before_batch", "of_which_counts+barrier"], list(arr)))

    def close(self):
        """Collective: every rank unmaps its peers before anybody frees."""
        if self.handle:
            lib().cldrd_node_detach(self.handle)
        if dist.is_initialized():
            try:
                dist.barrier(group=self.group)
            except Exception:
                pass
        if self.handle:
            lib().cldrd_node_destroy(self.handle)
            self.handle = C.c_void_p()
        self.ok = False


class _Lease:
    """Lives as long as a caller references the arrays of one result set; the set is handed out again afterwards."""

    def __init__(self, host: "_SharedHostResult", j: int):
        self.host, self.j = host, j
        host.busy[j] = True

    def __del__(self):
        try:
            self.host.busy[self.j] = False
        except Exception:
            pass


class _SharedHostResult:
    """Result sets of [rows, k] float32 scores + int64 ids in ONE shared-memory mapping that every rank of the node
    has page-locked (cldrd_host_register): every rank's merge kernel stores its slice of a result straight into the
    set over its own PCIe link, rank 0 hands the set to the caller as numpy arrays without any further copy.

    `index.search` returns arrays the caller owns, so a set goes back into rotation only when the caller
    has dropped every reference to its arrays (`_Lease`); rank 0 picks a free set per search and the choice reaches the
    other ranks' kernels through the exchange block (cldrd_node_search_begin_set).  The last set is scratch: when all
    others are still referenced the result is stored there and copied into fresh arrays."""

    _serial = 0
    ROTATION_MAX_BYTES = 256 << 20     # larger results (config 5: 1.2 GB) get one set + scratch

    def __init__(self, rank: int, rows: int, k: int, group, register: bool = True, device: int = 0):
        import mmap
        self.rank, self.group = rank, group
        self.cap_elems = rows * k
        self.stride = (self.cap_elems * 12 + 64 + 4095) // 4096 * 4096
        self.nsets = (3 if self.stride <= self.ROTATION_MAX_BYTES else 1) + 1
        self.busy = [False] * self.nsets
        self.nbytes = self.stride * self.nsets
        self.mm, self.base, self.ok = None, 0, False
        self.addr = 0            # host address of the mapping
        self.dev_base = 0        # the mapping as this rank's kernels address it
        # every step below ends in a collective that all ranks reach whatever failed locally
        name, fd = [None], -1
        if rank == 0:
            try:
                vfs = os.statvfs("/dev/shm")
                if vfs.f_bavail * vfs.f_frsize < self.nbytes + (16 << 20):
                    raise OSError("not enough room in /dev/shm")   # pinning would hit SIGBUS, not an error code
                _SharedHostResult._serial += 1
                path = f"/dev/shm/cldrd_{os.getpid()}_{_SharedHostResult._serial}"
                fd = os.open(path, os.O_CREAT | os.O_EXCL | os.O_RDWR, 0o600)
                name[0] = path
                os.ftruncate(fd, self.nbytes)
            except OSError:
                if fd >= 0:
                    os.close(fd)
                    fd = -1
                if name[0] is not None:
                    try:
                        os.unlink(name[0])
                    except OSError:
                        pass
                name[0] = None
        dist.broadcast_object_list(name, src=0, group=group)
        good = name[0] is not None
        if good:
            try:
                if rank != 0:
                    fd = os.open(name[0], os.O_RDWR)
                self.mm = mmap.mmap(fd, self.nbytes)
                self.addr = C.addressof(C.c_char.from_buffer(self.mm))
            except (OSError, ValueError):
                good = False
        if fd >= 0:
            os.close(fd)
        good = self._all_ok(good)         # also the barrier: everybody has mapped it (or given up)
        if rank == 0 and name[0] is not None:
            try:
                os.unlink(name[0])        # the mappings keep it alive; nothing is left behind on a crash
            except OSError:
                pass
        if good and register:
            self.base = self.addr
            reg = lib().cldrd_host_register(C.c_void_p(self.base), self.nbytes) == 0
            if not reg:
                self.base = 0
            else:
                dp = C.c_void_p()
                reg = lib().cldrd_host_device_ptr(device, C.c_void_p(self.base), C.byref(dp)) == 0
                self.dev_base = dp.value or 0
            good = self._all_ok(reg)
        self.ok = good
        if not good:
            self.close()

    def _all_ok(self, mine: bool) -> bool:
        return _all_ok(mine, self.group)

    def fits(self, rows: int, k: int) -> bool:
        return rows * k <= self.cap_elems

    def pick(self) -> int:
        """rank 0: a set no caller references any more, else the scratch set (the last one)."""
        for j in range(self.nsets - 1):
            if not self.busy[j]:
                return j
        return self.nsets - 1

    def set_ptrs(self, rows_alloc: int, k: int):
        """(scores, ids) device addresses of every set for a result of rows_alloc rows."""
        ioff = self.ids_offset(rows_alloc, k)
        return ([self.dev_base + j * self.stride for j in range(self.nsets)],
                [self.dev_base + j * self.stride + ioff for j in range(self.nsets)])

    def views(self, j: int, n: int, k: int, rows_alloc: int, lease: bool = True):
        """numpy arrays over set j: D [n,k] at byte 0, I [n,k] behind the score block of rows_alloc rows.  With
        lease=True the set stays out of rotation until both arrays (and every view of them) are gone."""
        import numpy as np
        buf = (C.c_char * self.stride).from_address(self.addr + j * self.stride)
        buf._keep = (self.mm, _Lease(self, j) if lease else None)
        D = np.frombuffer(buf, dtype=np.float32, count=n * k, offset=0).reshape(n, k)
        I = np.frombuffer(buf, dtype=np.int64, count=n * k, offset=self.ids_offset(rows_alloc, k)).reshape(n, k)
        return D, I

    @staticmethod
    def ids_offset(rows_alloc: int, k: int) -> int:
        return (rows_alloc * k * 4 + 63) // 64 * 64

    def close(self):
        if self.base:
            lib().cldrd_host_unregister(C.c_void_p(self.base))
            self.base = 0
            self.dev_base = 0

    def __del__(self):
        # the page-lock must go before the mapping does: a later mapping at the same address could not be
        # registered otherwise
        try:
            self.close()
        except Exception:
            pass


class ShardedSearcher:
    """rank-local shard + this rank's part of the sharded protocol.  Build with `from_rows` (device rows
    already in HBM, zero copy) or `from_file` (each rank preads only its own row range); `search` takes and
    returns CUDA tensors, `search_host` host arrays."""

    SEED_MIN_ROWS = 1 << 20   # below this the per-shard progressive scheme is already cheap
    RING = 3                  # batches of one search in flight (cldrd_node allows 4)

    def __init__(self, shard: _Shard, ntotal: int, d: int, id_map: Optional[torch.Tensor], group=None):
        self.shard = shard
        self.local = GpuIndexFlat(shard, shard.nrows, d)
        self.ntotal, self.d = ntotal, d
        self.id_map = id_map  # int64 [ntotal] on rank 0's device (replicated on first sharded use), or None
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self._nx: Optional[_NodeExchange] = None
        self._nx_disabled = False
        self._host: Optional[_SharedHostResult] = None
        self._host_disabled = False
        self._host_sets_for = None
        self._setup_done = False
        self._id_map_synced = False
        self.last_seed_misses = 0
        self.last_phase_ms = None

    @classmethod
    def from_rows(cls, rows: torch.Tensor, row0: int, ntotal: int, scan: str = "auto",
                  id_map: Optional[torch.Tensor] = None, group=None) -> "ShardedSearcher":
        assert rows.is_cuda and rows.dtype == torch.float32 and rows.is_contiguous()
        n, d = rows.shape
        sh = _Shard(rows.device.index, row0, n, d, scan)

        def filler(s: _Shard):
            check(lib().cldrd_shard_adopt(s.handle, C.c_void_p(rows.data_ptr())))

        sh.fill(filler)
        sh._keepalive = rows
        return cls(sh, ntotal, d, id_map, group)

    @classmethod
    def from_file(cls, path: str, device: int, scan: str = "auto", group=None) -> "ShardedSearcher":
        n, d = C.c_int64(), C.c_int32()
        has_ids = C.c_int32()
        check(lib().cldrd_index_probe(str(path).encode(), C.byref(n), C.byref(d), None, C.byref(has_ids), None,
                                      None, None))
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rr = shard_ranges(n.value, world)[rank]
        sh = _Shard(device, rr.start, len(rr), d.value, scan)

        def filler(s: _Shard):
            check(lib().cldrd_shard_load_file(s.handle, str(path).encode()))
            check(lib().cldrd_shard_set_ids(s.handle, None))  # ids are applied after the merge

        sh.fill(filler)
        id_map = None
        if has_ids.value and rank == 0:
            import numpy as np
            ids = np.empty((n.value,), dtype=np.int64)
            check(lib().cldrd_index_read_ids(str(path).encode(), 0, n.value, _lib.ptr(ids)))
            id_map = torch.from_numpy(ids).to(f"cuda:{device}")
        return cls(sh, n.value, d.value, id_map, group)

    # ---- one-time agreements between the shards of one index ---------------------------------------------

    def _setup(self):
        """The error band every decision of the protocol uses (filter cut, counted cut, seed check) must be the
        SAME on all shards: same scan precision (eps coefficient) and same row-norm bound.  scan="auto" picks f16
        or tf32 per shard from that shard's own value range, so the ranks settle on the coarsest one here."""
        if self.world == 1 or self._setup_done:
            return
        cdev = _coll_dev(self.group)
        order = {"f16": 0, "tf32": 1, "bf16": 2, "simt": 3}
        mode = torch.tensor([order[self.shard.scan], -order[self.shard.scan]], dtype=torch.int32, device=cdev)
        dist.all_reduce(mode, op=dist.ReduceOp.MAX, group=self.group)
        hi, lo = int(mode[0].item()), -int(mode[1].item())
        if hi != lo:
            if self.shard.scan_request != "auto" or hi != order["tf32"]:
                raise CldrdError(_lib.E_INVAL, f"the shards of one index use different scan modes ({lo} .. {hi}): "
                                               "pass the same scan= on every rank")
            if order[self.shard.scan] != hi:      # some other shard does not fit fp16: everybody scans in tf32
                self.shard.refill("tf32")
        b = C.c_float()
        check(lib().cldrd_shard_norm_bound(self.shard.handle, C.byref(b)))
        t = torch.tensor([b.value], dtype=torch.float32, device=cdev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        check(lib().cldrd_shard_set_norm_bound(self.shard.handle, C.c_float(float(t.item()))))
        self._setup_done = True

    def _id_map_everywhere(self):
        """Every rank's merge kernel applies the external ids to its slice: replicate rank 0's table once."""
        if self._id_map_synced:
            return
        dev = torch.device("cuda", self.shard.device)
        cdev = _coll_dev(self.group)
        has = torch.tensor([1 if (self.rank == 0 and self.id_map is not None) else 0], dtype=torch.int32, device=cdev)
        dist.broadcast(has, src=0, group=self.group)
        if int(has.item()):
            if self.rank != 0:
                self.id_map = torch.empty((self.ntotal,), dtype=torch.int64, device=dev)
            if cdev.type == "cuda":
                dist.broadcast(self.id_map, src=0, group=self.group)
            else:
                t = self.id_map.cpu()
                dist.broadcast(t, src=0, group=self.group)
                self.id_map.copy_(t)
        self._id_map_synced = True

    def _node(self, k: int) -> Optional[_NodeExchange]:
        """The node-local peer-memory exchange, or None when it is unavailable (CLDRD_DIST_P2P=0, IPC / peer
        access refused -- agreed by all ranks): then NCCL moves the lists."""
        if os.environ.get("CLDRD_DIST_P2P", "1") == "0" or self.world > _lib.MAX_PEERS or self._nx_disabled:
            return None
        nx = self._nx
        if nx is not None and k <= nx.max_k:
            return nx
        if nx is not None:
            nx.close()
        self._nx = _NodeExchange(self.shard.device, self.rank, self.world, int(k), self.group, self.d)
        if not self._nx.ok:
            self._nx = None
            self._nx_disabled = True
        return self._nx

    # ---- the node-wide search ------------------------------------------------------------------------------

    def _end_batch(self, nx: _NodeExchange, b0: int):
        nfail = C.c_int32()
        idx = (C.c_int32 * _lib.QUERY_BATCH)()
        check(lib().cldrd_node_search_end(self.shard.handle, nx.handle, C.byref(nfail), idx, _lib.QUERY_BATCH))
        return [b0 + idx[i] for i in range(nfail.value)]

    def _run_node(self, nx: _NodeExchange, q: torch.Tensor, k: int, seeded: bool, out_of_batch, out_rows=None,
                  after_batch=None, out_set=None, on_end=None, n_queries=None):
        """Queue the batches of one search (at most RING in flight), return the queries (indices into q) that
        have to be searched again -- the same list on every rank.  out_of_batch(b0, nb) -> (scores ptr, ids ptr)
        of the batch's output rows; out_rows: optional int32 device tensor, output row of every query.
        out_set (instead of out_of_batch): results go to a registered output set, rows from the batch's first
        query on; the value is the set rank 0 picked (>= 0 there, -1 on the ranks that follow).
        on_end(b0, nb, raised): called when a batch has ended (its rows have landed everywhere), in order.
        q: the replicated queries as a CUDA tensor, or a callable (b0, nb) -> device address of the batch's queries
        (with n_queries) that may enqueue whatever brings them there."""
        if callable(q):
            q_of_batch, n = q, int(n_queries)
        else:
            n = q.shape[0]

            def q_of_batch(b0, nb):
                return q[b0:b0 + nb].data_ptr()

        st = torch.cuda.current_stream(torch.device("cuda", self.shard.device)).cuda_stream
        idm = C.c_void_p(self.id_map.data_ptr()) if self.id_map is not None else None
        inflight, again = [], []

        def end_oldest():
            b0 = inflight.pop(0)
            raised = self._end_batch(nx, b0)
            again.extend(raised)
            if on_end is not None:
                on_end(b0, min(_lib.QUERY_BATCH, n - b0), raised)

        # several batches: the host runs ahead of the GPU anyway (RING batches queued), so the ranks sleep-poll instead of
        # spinning and leave their cores to whatever consumes the batches (the run-file writer's formatting threads)
        check(lib().cldrd_node_set_wait_mode(nx.handle, 1 if n > _lib.QUERY_BATCH else 0))
        t_host0 = time.perf_counter()
        with self.local._lock:
            for b0 in range(0, n, _lib.QUERY_BATCH):
                nb = min(_lib.QUERY_BATCH, n - b0)
                if len(inflight) >= self.RING:
                    end_oldest()
                rows = C.c_void_p(out_rows[b0:b0 + nb].data_ptr()) if out_rows is not None else None
                qp = q_of_batch(b0, nb)
                if out_set is not None:
                    check(lib().cldrd_node_search_begin_set(self.shard.handle, nx.handle, C.c_void_p(qp),
                                                            nb, int(k), 1 if seeded else 0, int(out_set),
                                                            0 if out_rows is not None else b0, rows, idm, C.c_void_p(st)))
                else:
                    oD, oI = out_of_batch(b0, nb)
                    check(lib().cldrd_node_search_begin(self.shard.handle, nx.handle, C.c_void_p(qp), nb,
                                                        int(k), 1 if seeded else 0, C.c_void_p(oD), C.c_void_p(oI), rows, idm,
                                                        C.c_void_p(st)))
                inflight.append(b0)
                if after_batch is not None:
                    after_batch(b0, nb)
            t_host1 = time.perf_counter()
            while inflight:
                end_oldest()
        self.last_phase_ms = nx.phase_ms()
        self.last_phase_ms["host_enqueue"] = (t_host1 - t_host0) * 1e3
        self.last_phase_ms["host_wait"] = (time.perf_counter() - t_host1) * 1e3
        return again

    def _check_q(self, q: torch.Tensor, k: int) -> torch.Tensor:
        if not (isinstance(q, torch.Tensor) and q.is_cuda and q.dtype == torch.float32 and q.dim() == 2 and
                q.shape[1] == self.d):
            raise TypeError(f"search: q must be a float32 CUDA tensor of shape [nq, {self.d}]")
        if not 1 <= int(k) <= _lib.MAX_K:
            raise RuntimeError(f"search: k={k} outside [1, {_lib.MAX_K}]")
        return q.contiguous()

    def search(self, q: torch.Tensor, k: int):
        """q: replicated float32 [nq,d] CUDA tensor.  Returns (D, I) on rank 0, (None, None) elsewhere."""
        q = self._check_q(q, k)
        k = int(k)
        if self.world == 1:
            D, I = self.local.search_device(q, k, translate_ids=False)
            if self.id_map is not None:
                return merge_candidates(D.unsqueeze(0), I.unsqueeze(0), self.id_map)
            return D, I
        n, dev = q.shape[0], q.device
        if n == 0:
            if self.rank != 0:
                return None, None
            return (torch.empty((0, k), dtype=torch.float32, device=dev), torch.empty((0, k), dtype=torch.int64, device=dev))
        self._setup()
        nx = self._node(k)
        if nx is None:
            return self._search_nccl(q, k)
        self._id_map_everywhere()
        st = torch.cuda.current_stream(dev).cuda_stream
        resD, resI = nx.result_ptrs(0)     # rank 0's result buffer, as this rank addresses it
        outD = outI = None
        if self.rank == 0:
            outD = torch.empty((n, k), dtype=torch.float32, device=dev)
            outI = torch.empty((n, k), dtype=torch.int64, device=dev)

        def collect(b0, nb):               # rank 0: batch rows out of the result buffer before the next batch lands
            if self.rank == 0:
                check(lib().cldrd_peer_copy(dev.index, C.c_void_p(outD[b0:].data_ptr()), C.c_void_p(resD), nb * k * 4, C.c_void_p(st)))
                check(lib().cldrd_peer_copy(dev.index, C.c_void_p(outI[b0:].data_ptr()), C.c_void_p(resI), nb * k * 8, C.c_void_p(st)))

        again = self._run_node(nx, q, k, self.ntotal >= self.SEED_MIN_ROWS, lambda b0, nb: (resD, resI), None, collect)
        self.last_seed_misses = len(again)
        if again:    # rare: the seed sat above the true k-th score, or a survivor buffer overflowed: unseeded retry
            idx = torch.tensor(again, dtype=torch.int64, device=dev)
            q2 = q[idx].contiguous()
            tmpD = torch.empty((len(again), k), dtype=torch.float32, device=dev) if self.rank == 0 else None
            tmpI = torch.empty((len(again), k), dtype=torch.int64, device=dev) if self.rank == 0 else None

            def collect2(b0, nb):
                if self.rank == 0:
                    check(lib().cldrd_peer_copy(dev.index, C.c_void_p(tmpD[b0:].data_ptr()), C.c_void_p(resD), nb * k * 4, C.c_void_p(st)))
                    check(lib().cldrd_peer_copy(dev.index, C.c_void_p(tmpI[b0:].data_ptr()), C.c_void_p(resI), nb * k * 8, C.c_void_p(st)))

            left = self._run_node(nx, q2, k, False, lambda b0, nb: (resD, resI), None, collect2)
            assert not left, "an unseeded batch cannot raise queries"
            if self.rank == 0:
                outD[idx] = tmpD
                outI[idx] = tmpI
        return (outD, outI) if self.rank == 0 else (None, None)

    def search_host(self, q_host, k: int, on_batch=None):
        """Host buffers in, host buffers out (the shape of the reference's `index.search(x, k)`,
        retriever/retrieval_utils.py:135).

        q_host: the replicated float32 [nq, d] queries in host memory (numpy array or CPU tensor; page-locked
        memory makes the upload asynchronous).  Returns numpy (D, I) on rank 0, (None, None) elsewhere; the arrays
        belong to the caller (they stay valid, untouched by later searches, for as long as they are referenced).
        On one node every rank's merge kernel stores its slice of the result straight into a shared page-locked
        result set over its own PCIe link; rank 0 hands out a set the caller no longer references."""
        qh = torch.as_tensor(q_host)
        assert qh.dtype == torch.float32 and qh.dim() == 2 and qh.shape[1] == self.d
        if not 1 <= int(k) <= _lib.MAX_K:
            raise RuntimeError(f"search: k={k} outside [1, {_lib.MAX_K}]")
        return self._to_host(None, qh.contiguous(), int(k), on_batch)

    def search_to_host(self, q: torch.Tensor, k: int, on_batch=None):
        """Device-resident queries in (e.g. straight from the encoder: SURVEY §8 f-3), host arrays out; otherwise
        `search_host`.  on_batch(b0, nb, D_rows, I_rows), rank 0 only: called in order as soon as the rows of an
        8192-query batch have landed in host memory, while later batches are still being searched -- the run-file
        writer of config 5 (502 939 queries) works on batch i while batch i+1 is scanned."""
        return self._to_host(self._check_q(q, k), None, int(k), on_batch)

    def _upload(self, qh: torch.Tensor) -> torch.Tensor:
        dev = torch.device("cuda", self.shard.device)
        n = qh.shape[0]
        stage = getattr(self, "_q_stage", None)
        if stage is None or stage.shape[0] < n:
            stage = self._q_stage = torch.empty((max(n, 1), self.d), dtype=torch.float32, device=dev)
        q = stage[:n]
        q.copy_(qh, non_blocking=True)
        return q

    def _to_host(self, q: Optional[torch.Tensor], qh: Optional[torch.Tensor], k: int, on_batch):
        """q: the queries on the device, or None with qh: the queries in host memory."""
        n = q.shape[0] if q is not None else qh.shape[0]
        dev = torch.device("cuda", self.shard.device)

        def via_device():
            D, I = self.search(q if q is not None else self._upload(qh), k)
            if self.rank != 0:
                return None, None
            D, I = D.cpu().numpy(), I.cpu().numpy()
            if on_batch is not None:
                for b0 in range(0, n, _lib.QUERY_BATCH):
                    on_batch(b0, min(_lib.QUERY_BATCH, n - b0), D[b0:b0 + _lib.QUERY_BATCH], I[b0:b0 + _lib.QUERY_BATCH])
            return D, I

        if self.world == 1 or n == 0:
            return via_device()
        self._setup()
        nx = self._node(k)
        if nx is None:          # no peer mapping: device result on rank 0, one copy down
            return via_device()
        self._id_map_everywhere()
        host = self._host
        if not self._host_disabled and (host is None or not host.fits(n, k)):
            if host is not None:
                host.close()
            host = self._host = _SharedHostResult(self.rank, n, k, self.group, device=dev.index)
            self._host_sets_for = None
            if not host.ok:           # agreed by all ranks
                host = self._host = None
                self._host_disabled = True
        if host is None:
            return via_device()
        if self._host_sets_for != (id(nx), n, k):      # the set addresses depend on the result shape
            sD, sI = host.set_ptrs(n, k)
            check(lib().cldrd_node_set_outputs(nx.handle, host.nsets, (C.c_void_p * host.nsets)(*sD),
                                               (C.c_void_p * host.nsets)(*sI)))
            self._host_sets_for = (id(nx), n, k)
        j = host.pick() if self.rank == 0 else -1
        Dv = Iv = None
        if self.rank == 0:
            Dv, Iv = host.views(j, n, k, n, lease=j != host.nsets - 1)
        held = []          # batches whose callback waits for the retry of a raised query (rare)
        scratch = j == host.nsets - 1

        def report(b0, nb):
            # rows of a leased set stay valid while the consumer holds them; the scratch set is reused by the next search
            Db, Ib = Dv[b0:b0 + nb], Iv[b0:b0 + nb]
            on_batch(b0, nb, Db.copy() if scratch else Db, Ib.copy() if scratch else Ib)

        def on_end(b0, nb, raised):
            if on_batch is None or self.rank != 0:
                return
            if raised or held:
                held.append((b0, nb))
            else:
                report(b0, nb)

        if q is None and nx.d == self.d:
            # Host queries: G uploads of the same 21 MB over G PCIe links pulling from one buffer cost more than the
            # whole exchange of the search.  Every rank uploads 1/G of the batch into its block and the part is stored
            # into every other rank's block over NVLink (cldrd_node_spread_queries).
            st = torch.cuda.current_stream(dev).cuda_stream
            qx, row_bytes, src = nx.query_ptr(), self.d * 4, qh.data_ptr()

            def q_src(b0, nb):
                sl = (nb + self.world - 1) // self.world
                lo = min(nb, self.rank * sl)
                hi = min(nb, lo + sl)
                if hi > lo:
                    check(lib().cldrd_peer_copy(dev.index, C.c_void_p(qx + lo * row_bytes), C.c_void_p(src + (b0 + lo) * row_bytes),
                                                (hi - lo) * row_bytes, C.c_void_p(st)))
                check(lib().cldrd_node_spread_queries(self.shard.handle, nx.handle, lo, hi - lo, C.c_void_p(st)))
                return qx
        else:
            if q is None:
                q = self._upload(qh)
            q_src = q
        again = self._run_node(nx, q_src, k, self.ntotal >= self.SEED_MIN_ROWS, None, out_set=j, on_end=on_end, n_queries=n)
        self.last_seed_misses = len(again)
        if again:
            rows = torch.tensor(again, dtype=torch.int32, device=dev)
            q2 = q[rows.long()].contiguous() if q is not None else qh[torch.tensor(again, dtype=torch.int64)].to(dev)
            left = self._run_node(nx, q2, k, False, None, out_rows=rows, out_set=j)
            assert not left, "an unseeded batch cannot raise queries"
        if self.rank != 0:
            return None, None
        for b0, nb in held:
            report(b0, nb)
        if scratch:                   # every rotating set is still referenced by the caller: scratch + one copy
            return Dv.copy(), Iv.copy()
        return Dv, Iv

    def close(self):
        """Collective: releases the exchange block, the shared host block and the shard."""
        if self._nx is not None:
            self._nx.close()
            self._nx = None
        if self._host is not None:
            self._host.close()
            self._host = None
        self.shard.close()

    # ---- NCCL transport (no peer mapping available) ------------------------------------------------------------

    def _trim(self, D: torch.Tensor, I: torch.Tensor):
        """A seeded shard returns far fewer than k valid rows per query (about 3.5k / world): agree
        on the widest valid prefix over all ranks (one scalar all-reduce) and gather only that."""
        if D.shape[0] == 0:
            return D, I
        w = (I >= 0).sum(dim=1).max().reshape(1)
        dist.all_reduce(w, op=dist.ReduceOp.MAX, group=self.group)
        w = max(1, min(D.shape[1], (int(w.item()) + 63) // 64 * 64))
        if w == D.shape[1]:
            return D, I
        return D[:, :w].contiguous(), I[:, :w].contiguous()

    def _search_nccl(self, q: torch.Tensor, k: int):
        """The same search with NCCL moving the data: all-gather of the sample scores -> seed; seeded search of
        the shard; all-to-all so that rank j receives every shard's lists for the j-th slice of the queries,
        merges and verifies that slice; only the merged [slice, k] results travel to rank 0 (gather)."""
        n = q.shape[0]
        prof = os.environ.get("CLDRD_DIST_PROFILE") == "1"
        marks = []

        def mark(name):
            if prof:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                marks.append((name, e))

        mark("start")
        seeded = self.ntotal >= self.SEED_MIN_ROWS and n > 0
        seed = None
        if seeded:
            topj = self.local.sample_device(q, k)
            allj = torch.empty((self.world,) + tuple(topj.shape), dtype=topj.dtype, device=topj.device)
            dist.all_gather_into_tensor(allj, topj, group=self.group)
            seed = torch.empty((n,), dtype=torch.float32, device=q.device)
            st = torch.cuda.current_stream(q.device).cuda_stream
            check(lib().cldrd_seed_from_samples(q.device.index, C.c_void_p(allj.data_ptr()), self.world, n,
                                                C.c_void_p(seed.data_ptr()), C.c_void_p(st)))
        mark("sample+allgather+seed")
        D, I, eps2 = self.local.search_device_seeded(q, k, seed)
        mark("seeded search")
        D, I = self._trim(D, I)
        world, W = self.world, D.shape[1]
        sl = (n + world - 1) // world
        n_pad = sl * world
        if n_pad != n:
            D = torch.cat([D, D.new_full((n_pad - n, W), -3.4028234663852886e38)])
            I = torch.cat([I, I.new_full((n_pad - n, W), -1)])
        recvD, recvI = torch.empty_like(D), torch.empty_like(I)
        dist.all_to_all_single(recvD, D, group=self.group)
        dist.all_to_all_single(recvI, I, group=self.group)
        mark("all-to-all")
        mD, mI = merge_candidates(recvD.view(world, sl, W), recvI.view(world, sl, W), None, k)
        nfail = torch.zeros((1,), dtype=torch.int64, device=q.device)
        fail_sl = torch.zeros((sl,), dtype=torch.int32, device=q.device)
        lo = self.rank * sl
        n_mine = max(0, min(sl, n - lo))
        if seeded and n_mine > 0:
            st = torch.cuda.current_stream(q.device).cuda_stream
            check(lib().cldrd_verify_seed(q.device.index, C.c_void_p(mD.data_ptr()), n_mine, k,
                                          C.c_void_p(seed[lo:lo + n_mine].contiguous().data_ptr()),
                                          C.c_void_p(eps2[lo:lo + n_mine].contiguous().data_ptr()),
                                          C.c_void_p(fail_sl.data_ptr()), C.c_void_p(st)))
            nfail[0] = fail_sl[:n_mine].sum()
        mark("merge+verify")
        outD = outI = None
        if self.rank == 0:
            allD = torch.empty((world, sl, k), dtype=mD.dtype, device=q.device)
            allI = torch.empty((world, sl, k), dtype=mI.dtype, device=q.device)
            dist.gather(mD, list(allD.unbind(0)), dst=0, group=self.group)
            dist.gather(mI, list(allI.unbind(0)), dst=0, group=self.group)
            outD, outI = allD.view(n_pad, k)[:n], allI.view(n_pad, k)[:n]
        else:
            dist.gather(mD, None, dst=0, group=self.group)
            dist.gather(mI, None, dst=0, group=self.group)
        mark("gather")
        self.last_seed_misses = 0
        if seeded:
            dist.all_reduce(nfail, op=dist.ReduceOp.SUM, group=self.group)
            if int(nfail.item()) > 0:   # rare: the seed sat above the true k-th score for these queries
                fail_all = torch.empty((world, sl), dtype=torch.int32, device=q.device)
                dist.all_gather_into_tensor(fail_all, fail_sl, group=self.group)
                idx = torch.nonzero(fail_all.view(-1)[:n]).flatten()
                D2, I2, _ = self.local.search_device_seeded(q[idx].contiguous(), k, None)
                D2, I2 = self._trim(D2, I2)
                allD2, allI2 = gather_candidates(D2, I2, dst=0, group=self.group)
                if self.rank == 0:
                    pD, pI = merge_candidates(allD2, allI2, None, k)
                    outD[idx] = pD
                    outI[idx] = pI
            self.last_seed_misses = int(nfail.item())
        mark("miss broadcast")
        if prof:
            torch.cuda.synchronize()
            self.last_phase_ms = {b[0]: a[1].elapsed_time(b[1]) for a, b in zip(marks, marks[1:])}
        if self.rank != 0:
            return None, None
        if self.id_map is not None:
            outI = torch.where(outI >= 0, self.id_map[outI.clamp_min(0)], outI)
        return outD, outI
