"""torchrun --nproc-per-node 2 tools/p2p_check.py : the peer-memory exchange alone (no search): every rank
fills its plane of every peer's buffer through the mapped pointers, then checks what the peers wrote."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cl-drd_b200")]
import torch
import torch.distributed as dist
from cldrd import dist as CD
from cldrd._lib import check, lib

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
for elems in (world * 36 * 100, world * 3490 * 1000):
    px = CD._PeerExchange(dev.index, rank, world, elems, None)
    print(f"rank {rank}: elems {elems} ok={px.ok} xD={[hex(p or 0) for p in px.xD]} oD={hex(px.oD or 0)}", flush=True)
    assert px.ok
    plane = elems // world
    src = torch.full((plane,), float(rank + 1), dtype=torch.float32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    for r in range(world):   # my plane in rank r's buffer
        check(lib().cldrd_peer_copy(dev.index, C.c_void_p(px.xD[r] + rank * plane * 4), C.c_void_p(src.data_ptr()), plane * 4, C.c_void_p(st)))
    check(lib().cldrd_peer_copy(dev.index, C.c_void_p(px.oD + rank * plane * 4), C.c_void_p(src.data_ptr()), plane * 4, C.c_void_p(st)))
    torch.cuda.synchronize()
    dist.barrier()
    got = torch.empty((elems,), dtype=torch.float32, device=dev)
    check(lib().cldrd_peer_copy(dev.index, C.c_void_p(got.data_ptr()), C.c_void_p(px.xD[rank]), elems * 4, C.c_void_p(st)))
    torch.cuda.synchronize()
    exp = torch.arange(1, world + 1, dtype=torch.float32, device=dev).repeat_interleave(plane)
    assert torch.equal(got, exp), (rank, got[::plane], exp[::plane])
    if rank == 0:
        check(lib().cldrd_peer_copy(dev.index, C.c_void_p(got.data_ptr()), C.c_void_p(px.oD), elems * 4, C.c_void_p(st)))
        torch.cuda.synchronize()
        assert torch.equal(got, exp)
    dist.barrier()
    # kernel-level peer access: the merge kernel reads my (local) planes and stores into rank 0's buffer
    k = 100
    sl = elems // world // k
    xI_fill = torch.arange(elems, dtype=torch.int64, device=dev)
    check(lib().cldrd_peer_copy(dev.index, C.c_void_p(px.xI[rank]), C.c_void_p(xI_fill.data_ptr()), elems * 8, C.c_void_p(st)))
    torch.cuda.synchronize()
    lo = rank * sl
    print(f"rank {rank}: merge kernel -> rank 0 buffer, sl={sl}", flush=True)
    check(lib().cldrd_merge_w(dev.index, C.c_void_p(px.xD[rank]), C.c_void_p(px.xI[rank]), world, sl, k, k, None,
                              C.c_void_p(px.oD + lo * k * 4), C.c_void_p(px.oI + lo * k * 8), C.c_void_p(st)))
    torch.cuda.synchronize()
    print(f"rank {rank}: merge kernel done", flush=True)
    dist.barrier()
    px.close()
print(f"rank {rank}: p2p exchange ok", flush=True)
dist.destroy_process_group()
