"""World-size-2 gloo test (CPU) of the sharding plumbing: shard ranges, the candidate gather, and
that merging per-shard oracle results in key order reproduces the whole-index oracle result."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "cl-drd_b200")]
    from cldrd import dist as CD
    from cldrd.index import shard_ranges
    from oracle import flat_ip as O
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    xb, xq = O.synth(5001, 32, 0), O.synth(17, 32, 1)
    rr = shard_ranges(5001, world)[rank]
    D, R = O.search_rows(xb[rr.start:rr.stop], xq, 50)        # stand-in for the shard-local GPU search
    R = np.where(R >= 0, R + rr.start, -1)
    allD, allI = CD.gather_candidates(torch.from_numpy(D), torch.from_numpy(R), dst=0)
    if rank == 0:
        assert allD.shape == (world, 17, 50)
        d = allD.permute(1, 0, 2).reshape(17, -1).numpy()
        i = allI.permute(1, 0, 2).reshape(17, -1).numpy()
        order = np.lexsort((i, -d), axis=1)[:, :50]            # score desc, row asc == the merge kernel's key order
        Dm, Im = np.take_along_axis(d, order, 1), np.take_along_axis(i, order, 1)
        D_ref, R_ref = O.search_rows(xb, xq, 50)
        torch.save({"ok": bool(np.array_equal(Im, R_ref) and np.array_equal(Dm, D_ref))}, out)
    else:
        assert allD is None and allI is None
    dist.barrier()
    dist.destroy_process_group()


def test_gather_candidates_gloo_world2(tmp_path):
    out = str(tmp_path / "res.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert torch.load(out)["ok"]


def _shared_block_worker(rank, world, port, out):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "cl-drd_b200")]
    from cldrd import dist as CD
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    res = {}
    before = set(os.listdir("/dev/shm"))
    # (1) the mapping is really shared: every rank writes its slice, rank 0 reads all of them
    n, k, sl = 7, 5, 4                                            # 7 queries, slices of 4 rows, 2 ranks
    blk = CD._SharedHostResult(rank, world * sl, k, None, register=False)
    res["mapped"] = blk.ok
    assert blk.nsets == 4 and blk.pick() == 0
    D, I = blk.views(0, world * sl, k, world * sl)
    assert blk.pick() == 1                      # set 0 is referenced by D and I: the next result goes to set 1
    lo = rank * sl
    D[lo:lo + sl] = float(rank + 1)
    I[lo:lo + sl] = 100 * (rank + 1) + np.arange(k)
    dist.barrier()
    if rank == 0:
        res["shared"] = bool((D[:sl] == 1).all() and (D[sl:] == 2).all() and (I[sl:, 0] == 200).all())
        res["ids_aligned"] = CD._SharedHostResult.ids_offset(world * sl, k) % 8 == 0 and I.ctypes.data % 8 == 0
    view = D[1:3]
    del D, I
    res["view_keeps_the_set"] = blk.pick() == 1     # a slice of a handed-out array still owns the set
    del view
    res["set_returns_when_dropped"] = blk.pick() == 0
    dist.barrier()
    # (2) without CUDA the page-lock is refused: all ranks must agree on ok == False, nobody hangs
    blk2 = CD._SharedHostResult(rank, world * sl, k, None, register=True)
    res["register_refused_everywhere"] = (not blk2.ok) and blk2.base == 0
    res["no_files_left"] = set(os.listdir("/dev/shm")) <= before
    if rank == 0:
        torch.save(res, out)
    dist.barrier()
    dist.destroy_process_group()


def test_shared_host_result_block_gloo_world2(tmp_path):
    """cldrd.dist._SharedHostResult (the block ShardedSearcher.search_host writes its slices into): shared
    between the ranks, unlinked at once, and a refused cudaHostRegister (no GPU here) is agreed on collectively."""
    if torch.cuda.is_available():
        import pytest
        pytest.skip("the refusal path needs a box without CUDA")
    out = str(tmp_path / "res.pt")
    mp.spawn(_shared_block_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    res = torch.load(out)
    assert all(res.values()), res


def test_shard_ranges_cover_rows_exactly():
    sys.path[:0] = [os.path.join(ROOT, "cl-drd_b200")]
    from cldrd.index import shard_ranges
    for n, g in [(8841823, 8), (10, 3), (5, 8), (0, 2)]:
        rr = shard_ranges(n, g)
        assert rr[0].start == 0 and rr[-1].stop == n
        assert all(a.stop == b.start for a, b in zip(rr, rr[1:]))
        assert max(len(r) for r in rr) - min(len(r) for r in rr) <= 1
    assert [len(r) for r in shard_ranges(8841823, 8)].count(1105228) == 7
