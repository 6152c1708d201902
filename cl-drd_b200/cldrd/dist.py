"""Row-sharded search with one process per GPU (torchrun), the B200-native form of the
reference's `index_cpu_to_gpu_multiple(..., shard=True)` (retriever/retrieval_utils.py:174-182).

rank r keeps passage rows shard_ranges(N, G)[r] resident in its HBM; queries are replicated.
Per query batch: every rank scans a small sample of its rows, the sample scores are
all-gathered (NCCL) and turned into 32 levels per query, the last of which is the global filter seed;
every rank runs the fused scan + filter over its shard with that seed and counts its candidates above
each level; an all-reduce of those counts tells every rank which candidates can still reach the global
top-k, and only those are re-scored in fp32.  The re-score kernel stores each query's list straight
into the HBM of the rank that merges that query (peer memory over NVLink, mapped with CUDA IPC); every
rank merges and verifies its slice of the queries (same u64 key order as the single-GPU search, so
results are bit-identical) and stores it into rank 0's result buffer, where the id_map gather finishes
the job.  Without peer mapping an NCCL all-to-all and a gather move the lists.  torch.distributed is
plumbing only; no collective touches the index.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Tuple

import torch
import torch.distributed as dist

from . import _lib
from ._lib import check, lib
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


class _PeerExchange:
    """Exchange buffers of one node's ranks, mapped into every rank (CUDA IPC over NVLink / NVSwitch).

    Every rank owns  xD float32 [world][slice][k]  and  xI int64 [world][slice][k]  (plane p is written
    by rank p's re-score kernel: `cldrd_search_dev_scatter`); rank 0 also owns the result buffers
    oD / oI [world*slice][k] that every rank's merge kernel writes its slice into.  Sized for
    `cap_elems` = world*slice*k entries; rebuilt (collectively) when a search needs more."""

    def __init__(self, device: int, rank: int, world: int, cap_elems: int, group):
        self.device, self.rank, self.world, self.group = device, rank, world, group
        self.cap_elems = cap_elems
        self.own = []          # pointers from cldrd_peer_alloc
        self.opened = []       # pointers from cldrd_peer_open
        self.xD = [None] * world
        self.xI = [None] * world
        self.oD = self.oI = None
        elems = cap_elems
        sizes = [elems * 4, elems * 8] + ([elems * 4, elems * 8] if rank == 0 else [])
        handles = torch.zeros((4, _lib.PEER_HANDLE_BYTES), dtype=torch.uint8)
        ok = True
        try:
            for i, nbytes in enumerate(sizes):
                ptr = C.c_void_p()
                buf = (C.c_ubyte * _lib.PEER_HANDLE_BYTES)()
                check(lib().cldrd_peer_alloc(device, nbytes, C.byref(ptr), buf))
                self.own.append(ptr.value)
                handles[i] = torch.frombuffer(bytearray(buf), dtype=torch.uint8)
        except Exception:
            ok = False
        dev = torch.device("cuda", device)
        allh = torch.empty((world, 4, _lib.PEER_HANDLE_BYTES), dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(allh, handles.to(dev), group=group)
        allh = allh.cpu()
        if ok:
            try:
                for r in range(world):
                    for i in range(4 if r == 0 else 2):
                        if r == rank:
                            ptr_v = self.own[i]
                        else:
                            ptr = C.c_void_p()
                            hb = (C.c_ubyte * _lib.PEER_HANDLE_BYTES).from_buffer_copy(bytes(allh[r, i].tolist()))
                            check(lib().cldrd_peer_open(device, hb, C.byref(ptr)))
                            self.opened.append(ptr.value)
                            ptr_v = ptr.value
                        if i == 0:
                            self.xD[r] = ptr_v
                        elif i == 1:
                            self.xI[r] = ptr_v
                        elif i == 2:
                            self.oD = ptr_v
                        else:
                            self.oI = ptr_v
            except Exception:
                ok = False
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        self.ok = bool(flag.item())
        if self.ok:
            self.c_xD = (C.c_void_p * world)(*self.xD)
            self.c_xI = (C.c_void_p * world)(*self.xI)
        else:
            self.close()

    def fits(self, slice_rows: int, k: int) -> bool:
        return self.ok and self.world * slice_rows * k <= self.cap_elems

    def close(self):
        # every rank unmaps before anybody frees
        for p in self.opened:
            lib().cldrd_peer_close(self.device, C.c_void_p(p))
        self.opened = []
        if dist.is_initialized():
            try:
                torch.cuda.synchronize(self.device)
                dist.barrier(group=self.group)
            except Exception:
                pass
        for p in self.own:
            lib().cldrd_peer_free(self.device, C.c_void_p(p))
        self.own = []
        self.ok = False


class _SharedHostResult:
    """[rows, k] float32 scores + int64 ids in ONE shared-memory mapping that every rank of the node has
    page-locked (cldrd_host_register): each rank copies its merged slice device -> host over its own PCIe
    link, rank 0 reads the whole result as numpy arrays without any further copy."""

    _serial = 0

    def __init__(self, rank: int, rows: int, k: int, group, register: bool = True):
        import mmap
        self.rank, self.group = rank, group
        self.cap_elems = rows * k
        self.nbytes = self.cap_elems * 12 + 64
        self.mm, self.base, self.ok = None, 0, False
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
            self.base = C.addressof(C.c_char.from_buffer(self.mm))
            reg = lib().cldrd_host_register(C.c_void_p(self.base), self.nbytes) == 0
            if not reg:
                self.base = 0
            good = self._all_ok(reg)
        self.ok = good
        if not good:
            self.close()

    def _all_ok(self, mine: bool) -> bool:
        on_gpu = dist.get_backend(self.group) == "nccl"
        flag = torch.tensor([1 if mine else 0], dtype=torch.int32,
                            device=torch.device("cuda", torch.cuda.current_device()) if on_gpu else "cpu")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        return bool(flag.item())

    def fits(self, rows: int, k: int) -> bool:
        return rows * k <= self.cap_elems

    def views(self, n: int, k: int, rows_alloc: int):
        """numpy views: D [n,k] at byte 0, I [n,k] behind the score block of rows_alloc rows."""
        import numpy as np
        D = np.frombuffer(self.mm, dtype=np.float32, count=n * k, offset=0).reshape(n, k)
        I = np.frombuffer(self.mm, dtype=np.int64, count=n * k, offset=self.ids_offset(rows_alloc, k)).reshape(n, k)
        return D, I

    @staticmethod
    def ids_offset(rows_alloc: int, k: int) -> int:
        return (rows_alloc * k * 4 + 63) // 64 * 64

    def close(self):
        if self.base:
            lib().cldrd_host_unregister(C.c_void_p(self.base))
            self.base = 0

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

    def __init__(self, shard: _Shard, ntotal: int, d: int, id_map: Optional[torch.Tensor], group=None):
        self.shard = shard
        self.local = GpuIndexFlat(shard, shard.nrows, d)
        self.ntotal, self.d = ntotal, d
        self.id_map = id_map  # int64 [ntotal] on rank 0's device, or None
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1

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

    SEED_MIN_ROWS = 1 << 20   # below this the per-shard progressive scheme is already cheap

    def _sync_norm_bound(self):
        """Shards of one index must use the same row-norm bound in their error band."""
        if self.world == 1 or getattr(self, "_norm_synced", False):
            return
        b = C.c_float()
        check(lib().cldrd_shard_norm_bound(self.shard.handle, C.byref(b)))
        t = torch.tensor([b.value], dtype=torch.float32, device=f"cuda:{self.shard.device}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        check(lib().cldrd_shard_set_norm_bound(self.shard.handle, C.c_float(float(t.item()))))
        self._norm_synced = True

    def _levels(self, q: torch.Tensor, k: int) -> torch.Tensor:
        """Steps 1+2: sample every shard, all-gather the per-shard sample scores (NCCL), keep the SEED_J best
        of the union per query, best first.  The last column is the seed (a scan-score threshold near rank
        3.5k of the whole index); the columns before it are the levels the shards count against."""
        n = q.shape[0]
        topj = self.local.sample_device(q, k)
        allj = torch.empty((self.world,) + tuple(topj.shape), dtype=topj.dtype, device=topj.device)
        dist.all_gather_into_tensor(allj, topj, group=self.group)
        levels = torch.empty((n, _lib.SEED_J), dtype=torch.float32, device=q.device)
        st = torch.cuda.current_stream(q.device).cuda_stream
        check(lib().cldrd_levels_from_samples(q.device.index, C.c_void_p(allj.data_ptr()), self.world, n,
                                              C.c_void_p(levels.data_ptr()), C.c_void_p(st)))
        return levels

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

    def _peer_exchange(self, n: int, k: int):
        """The node-local peer-memory exchange for an [n, k] search, or None when it is unavailable
        (CLDRD_DIST_P2P=0, CPU process group, IPC / peer access refused): then NCCL moves the lists."""
        if os.environ.get("CLDRD_DIST_P2P", "1") == "0" or self.world > _lib.MAX_PEERS or n == 0:
            return None
        if getattr(self, "_px_disabled", False):
            return None
        sl = (n + self.world - 1) // self.world
        px = getattr(self, "_px", None)
        if px is not None and px.fits(sl, k):
            return px
        cap = self.world * sl * k
        if px is not None:
            cap = max(cap, px.cap_elems)
            px.close()
        self._px = _PeerExchange(self.shard.device, self.rank, self.world, cap, self.group)
        if not self._px.ok:
            self._px = None
            self._px_disabled = True   # agreed by all ranks (all-reduce MIN): nobody retries
        return self._px

    def _search_p2p(self, px: "_PeerExchange", q: torch.Tensor, k: int, levels, seed, mark, marks, prof, host=None):
        """Per batch of <= 8192 queries: scan + select with the seed and count the candidates above every
        sample level; all-reduce the counts (32 ints per query) so that every shard knows a threshold that k
        rows of the WHOLE index reach, and re-scores only what lies above it (about k / world rows per query);
        the re-score kernel stores every query's list in the merging rank's memory (NVLink peer stores).
        Then slice-wise merge + verify on every rank, the merge kernels storing into rank 0's result buffer.
        Two scalar all-reduces are the barriers between the kernels' peer accesses."""
        n, world, dev = q.shape[0], self.world, q.device
        seeded = levels is not None
        sl = (n + world - 1) // world
        st = torch.cuda.current_stream(dev).cuda_stream
        eps2 = torch.empty((n,), dtype=torch.float32, device=dev)
        token = torch.zeros((1,), dtype=torch.int32, device=dev)
        with self.local._lock:
            for b0 in range(0, n, _lib.QUERY_BATCH):
                nb = min(_lib.QUERY_BATCH, n - b0)
                qb = q[b0:b0 + nb]
                lv = levels[b0:b0 + nb] if seeded else None
                counts = torch.empty((nb, _lib.SEED_J), dtype=torch.int32, device=dev)
                check(lib().cldrd_scatter_begin(self.shard.handle, C.c_void_p(qb.data_ptr()), nb, int(k),
                                                C.c_void_p(lv.data_ptr()) if seeded else None,
                                                C.c_void_p(counts.data_ptr()), C.c_void_p(eps2[b0:].data_ptr()),
                                                C.c_void_p(st)))
                if seeded:
                    dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=self.group)
                check(lib().cldrd_scatter_finish(self.shard.handle, C.c_void_p(counts.data_ptr()),
                                                 C.c_void_p(lv.data_ptr()) if seeded else None, world, self.rank, sl, b0,
                                                 px.c_xD, px.c_xI, C.c_void_p(st)))
        mark("seeded search + scatter")
        dist.all_reduce(token, group=self.group)            # every plane of my buffers is written
        if os.environ.get("CLDRD_DIST_DEBUG") == "1":
            torch.cuda.synchronize()
            print(f"[rank {self.rank}] scatter + barrier ok: n={n} k={k} sl={sl} cap={px.cap_elems}", flush=True)
        lo = self.rank * sl
        n_mine = max(0, min(sl, n - lo))
        nfail = torch.zeros((1,), dtype=torch.int64, device=dev)
        fail_sl = torch.zeros((sl,), dtype=torch.int32, device=dev)
        if n_mine > 0:
            if host is None:      # merged slice -> rank 0's device buffer (peer stores by the merge kernel)
                outD = px.oD + lo * k * 4
                outI = px.oI + lo * k * 8
                idm = None
            else:                 # merged slice (external ids applied) stays here, then goes to the shared host block
                mD = torch.empty((n_mine, k), dtype=torch.float32, device=dev)
                mI = torch.empty((n_mine, k), dtype=torch.int64, device=dev)
                outD, outI = mD.data_ptr(), mI.data_ptr()
                idm = C.c_void_p(self.id_map.data_ptr()) if self.id_map is not None else None
            check(lib().cldrd_merge_planes(dev.index, C.c_void_p(px.xD[self.rank]), C.c_void_p(px.xI[self.rank]), world, sl,
                                           n_mine, k, k, idm, C.c_void_p(outD), C.c_void_p(outI), C.c_void_p(st)))
            if host is not None:
                rows_alloc = world * sl
                check(lib().cldrd_peer_copy(dev.index, C.c_void_p(host.base + lo * k * 4), C.c_void_p(outD), n_mine * k * 4,
                                            C.c_void_p(st)))
                check(lib().cldrd_peer_copy(dev.index, C.c_void_p(host.base + host.ids_offset(rows_alloc, k) + lo * k * 8), C.c_void_p(outI),
                                            n_mine * k * 8, C.c_void_p(st)))
            if seeded:
                check(lib().cldrd_verify_seed(dev.index, C.c_void_p(outD), n_mine, k,
                                              C.c_void_p(seed[lo:lo + n_mine].contiguous().data_ptr()),
                                              C.c_void_p(eps2[lo:lo + n_mine].contiguous().data_ptr()),
                                              C.c_void_p(fail_sl.data_ptr()), C.c_void_p(st)))
                nfail[0] = fail_sl[:n_mine].sum()
        mark("merge+verify")
        dist.all_reduce(nfail, op=dist.ReduceOp.SUM, group=self.group)   # also: every slice is in rank 0's buffer
        if host is not None:
            return self._finish_host(host, q, k, n, sl, seeded, nfail, fail_sl, mark, marks, prof)
        outD_t = outI_t = None
        if self.rank == 0:
            outD_t = torch.empty((n, k), dtype=torch.float32, device=dev)
            outI_t = torch.empty((n, k), dtype=torch.int64, device=dev)
            check(lib().cldrd_peer_copy(dev.index, C.c_void_p(outD_t.data_ptr()), C.c_void_p(px.oD), n * k * 4, C.c_void_p(st)))
            check(lib().cldrd_peer_copy(dev.index, C.c_void_p(outI_t.data_ptr()), C.c_void_p(px.oI), n * k * 8, C.c_void_p(st)))
        mark("result")
        misses = int(nfail.item()) if seeded else 0
        if misses > 0:   # rare: the seed sat above the true k-th score for these queries
            fail_all = torch.empty((world, sl), dtype=torch.int32, device=dev)
            dist.all_gather_into_tensor(fail_all, fail_sl, group=self.group)
            idx = torch.nonzero(fail_all.view(-1)[:n]).flatten()
            D2, I2, _ = self.local.search_device_seeded(q[idx].contiguous(), k, None)
            D2, I2 = self._trim(D2, I2)
            allD2, allI2 = gather_candidates(D2, I2, dst=0, group=self.group)
            if self.rank == 0:
                pD, pI = merge_candidates(allD2, allI2, None, k)
                outD_t[idx] = pD
                outI_t[idx] = pI
        self.last_seed_misses = misses
        mark("miss broadcast")
        if prof:
            torch.cuda.synchronize()
            self.last_phase_ms = {b[0]: a[1].elapsed_time(b[1]) for a, b in zip(marks, marks[1:])}
        if self.rank != 0:
            return None, None
        if self.id_map is not None:
            outI_t = torch.where(outI_t >= 0, self.id_map[outI_t.clamp_min(0)], outI_t)
        return outD_t, outI_t

    def _finish_host(self, host, q, k, n, sl, seeded, nfail, fail_sl, mark, marks, prof):
        """Tail of the host-result search: every rank's slice is in the shared block once the miss-count
        all-reduce (already issued, stream-ordered behind the copies) has completed."""
        world, dev = self.world, q.device
        misses = int(nfail.item()) if seeded else 0      # synchronises this rank's stream
        if not seeded:
            torch.cuda.current_stream(dev).synchronize()
        D = I = None
        if self.rank == 0:
            D, I = host.views(n, k, world * sl)
        mark("result")
        if misses > 0:
            fail_all = torch.empty((world, sl), dtype=torch.int32, device=dev)
            dist.all_gather_into_tensor(fail_all, fail_sl, group=self.group)
            idx = torch.nonzero(fail_all.view(-1)[:n]).flatten()
            D2, I2, _ = self.local.search_device_seeded(q[idx].contiguous(), k, None)
            D2, I2 = self._trim(D2, I2)
            allD2, allI2 = gather_candidates(D2, I2, dst=0, group=self.group)
            if self.rank == 0:
                pD, pI = merge_candidates(allD2, allI2, self.id_map, k)
                rows = idx.cpu().numpy()
                D[rows] = pD.cpu().numpy()
                I[rows] = pI.cpu().numpy()
        self.last_seed_misses = misses
        mark("miss broadcast")
        if prof:
            torch.cuda.synchronize()
            self.last_phase_ms = {b[0]: a[1].elapsed_time(b[1]) for a, b in zip(marks, marks[1:])}
        return D, I

    def _id_map_everywhere(self):
        """The host-result path applies external ids inside every rank's merge: replicate rank 0's table once."""
        if getattr(self, "_id_map_synced", False):
            return
        dev = torch.device("cuda", self.shard.device)
        has = torch.tensor([1 if (self.rank == 0 and self.id_map is not None) else 0], dtype=torch.int32, device=dev)
        dist.broadcast(has, src=0, group=self.group)
        if int(has.item()):
            if self.rank != 0:
                self.id_map = torch.empty((self.ntotal,), dtype=torch.int64, device=dev)
            dist.broadcast(self.id_map, src=0, group=self.group)
        self._id_map_synced = True

    def search_host(self, q_host, k: int):
        """Host buffers in, host buffers out (the shape of the reference's `index.search(x, k)`).

        q_host: the replicated float32 [nq, d] queries in host memory (numpy array or CPU tensor; page-locked
        memory makes the upload asynchronous).  Returns numpy (D, I) on rank 0, (None, None) elsewhere.  On one
        node the ranks write their slices of the result into one shared page-locked block, each over its own
        PCIe link, and rank 0 returns views of that block: they are valid until the next search_host call."""
        qh = torch.as_tensor(q_host)
        assert qh.dtype == torch.float32 and qh.dim() == 2 and qh.shape[1] == self.d
        n = qh.shape[0]
        dev = torch.device("cuda", self.shard.device)
        stage = getattr(self, "_q_stage", None)
        if stage is None or stage.shape[0] < n:
            stage = self._q_stage = torch.empty((max(n, 1), self.d), dtype=torch.float32, device=dev)
        q = stage[:n]
        q.copy_(qh, non_blocking=True)
        if self.world == 1 or n == 0:
            D, I = self.search(q, k)
            return (D.cpu().numpy(), I.cpu().numpy()) if self.rank == 0 else (None, None)
        self._sync_norm_bound()
        px = self._peer_exchange(n, k)
        if px is None:          # no peer mapping: device result on rank 0, one copy down
            D, I = self.search(q, k)
            return (D.cpu().numpy(), I.cpu().numpy()) if self.rank == 0 else (None, None)
        self._id_map_everywhere()
        sl = (n + self.world - 1) // self.world
        host = getattr(self, "_host", None)
        if not getattr(self, "_host_disabled", False) and (host is None or not host.fits(self.world * sl, k)):
            if host is not None:
                host.close()
            host = self._host = _SharedHostResult(self.rank, self.world * sl, k, self.group)
            if not host.ok:           # agreed by all ranks
                host = self._host = None
                self._host_disabled = True
        if host is None:
            D, I = self.search(q, k)
            return (D.cpu().numpy(), I.cpu().numpy()) if self.rank == 0 else (None, None)
        prof = os.environ.get("CLDRD_DIST_PROFILE") == "1"
        marks = []

        def mark(name):
            if prof:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                marks.append((name, e))

        mark("start")
        seeded = self.ntotal >= self.SEED_MIN_ROWS
        levels = self._levels(q, k) if seeded else None
        seed = levels[:, _lib.SEED_J - 1].contiguous() if seeded else None
        mark("sample+allgather+seed")
        return self._search_p2p(px, q, k, levels, seed, mark, marks, prof, host=host)

    def search(self, q: torch.Tensor, k: int):
        """q: replicated float32 [nq,d] CUDA tensor.  Returns (D, I) on rank 0, (None, None) elsewhere."""
        if self.world == 1:
            D, I = self.local.search_device(q, k, translate_ids=False)
            if self.id_map is not None:
                return merge_candidates(D.unsqueeze(0), I.unsqueeze(0), self.id_map)
            return D, I
        self._sync_norm_bound()
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
        levels = self._levels(q, k) if seeded else None
        seed = levels[:, _lib.SEED_J - 1].contiguous() if seeded else None
        mark("sample+allgather+seed")
        px = self._peer_exchange(n, k)
        if px is not None:
            return self._search_p2p(px, q, k, levels, seed, mark, marks, prof)
        D, I, eps2 = self.local.search_device_seeded(q, k, seed)
        mark("seeded search")
        D, I = self._trim(D, I)
        # Distributed merge: all-to-all so that rank j receives every shard's lists for the j-th
        # slice of the queries, merges and verifies that slice, and only the merged [slice, k]
        # results travel to rank 0.  (A plain gather would make rank 0 receive and merge everything.)
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
