"""torchrun worker for test_torchrun_two_ranks_nccl: every rank holds a row shard; rank 0 checks the
merged result against a single-GPU search of the whole index and against the oracle."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cl-drd_b200")]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from cldrd import dist as CD
    from cldrd.index import shard_ranges
    from oracle import flat_ip as O
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=dev)
    res = {}
    for name, n, d, nq, k in [("small", 50_000, 64, 70, 100), ("big", 1_300_000, 64, 70, 100)]:
        rng = np.random.Generator(np.random.PCG64(300))
        xb = rng.standard_normal((n, d), dtype=np.float32)
        xq = rng.standard_normal((nq, d), dtype=np.float32)
        ids = O.synth_ids(n, 301)
        rr = shard_ranges(n, world)[rank]
        rows = torch.from_numpy(xb[rr.start:rr.stop]).to(dev)
        q = torch.from_numpy(xq).to(dev)
        id_map = torch.from_numpy(ids).to(dev) if rank == 0 else None
        s = CD.ShardedSearcher.from_rows(rows, rr.start, n, scan="f16", id_map=id_map)
        D, I = s.search(q, k)
        if rank == 0:
            full = torch.from_numpy(xb).to(dev)
            one = CD.ShardedSearcher.from_rows(full, 0, n, scan="f16", id_map=id_map)
            one.world, one.rank = 1, 0
            D1, I1 = one.search(q, k)
            res[f"bit_equal_{name}"] = bool(torch.equal(D, D1) and torch.equal(I, I1))
            if name == "big":
                D_ref, I_ref = O.search(xb, ids, xq, k)
                r = O.compare_topk(D.cpu().numpy(), I.cpu().numpy(), D_ref, I_ref, *O.search(xb, ids, xq, k + 16, dtype=np.float64))
                res["oracle_ok"] = bool(r["ok"])
                res["seed_misses"] = getattr(s, "last_seed_misses", None)
        dist.barrier()
    if rank == 0:
        with open(args.out, "w") as f:
            json.dump(res, f)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
