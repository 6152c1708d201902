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
    ap.add_argument("--backend", default="nccl")
    ap.add_argument("--same-device", action="store_true", help="every rank on cuda:0 (one-GPU box; needs gloo)")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from cldrd import dist as CD
    from cldrd.index import shard_ranges
    from oracle import flat_ip as O
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = 0 if args.same_device else int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if args.backend == "nccl":
        dist.init_process_group("nccl", device_id=dev)
    else:
        dist.init_process_group(args.backend)
    transports = ("p2p", "nccl") if args.backend == "nccl" else ("p2p",)
    res = {}
    # (name, rows, d, queries, k, seed bias): "manyq" crosses the 8192-query batch of the engine and is not a
    # multiple of the world size; "miss" pushes every seed above every score so that all queries take the retry
    cases = [("small", 50_000, 64, 71, 100, None), ("big", 1_300_000, 64, 71, 100, None),
             ("manyq", 60_000, 64, 8301, 10, None), ("miss", 1_300_000, 64, 33, 50, "1e6")]
    for name, n, d, nq, k, bias in cases:
        rng = np.random.Generator(np.random.PCG64(300))
        xb = rng.standard_normal((n, d), dtype=np.float32)
        xq = rng.standard_normal((nq, d), dtype=np.float32)
        ids = O.synth_ids(n, 301)
        rr = shard_ranges(n, world)[rank]
        rows = torch.from_numpy(xb[rr.start:rr.stop]).to(dev)
        q = torch.from_numpy(xq).to(dev)
        id_map = torch.from_numpy(ids).to(dev) if rank == 0 else None
        os.environ["CLDRD_SEED_BIAS"] = bias or "0"
        s = CD.ShardedSearcher.from_rows(rows, rr.start, n, scan="f16", id_map=id_map)
        os.environ["CLDRD_SEED_BIAS"] = "0"
        out = {}
        for transport in transports:           # peer-memory exchange (default) and the NCCL all-to-all path
            os.environ["CLDRD_DIST_P2P"] = "1" if transport == "p2p" else "0"
            out[transport] = s.search(q, k)
            torch.cuda.synchronize()
            print(f"[rank {rank}] {name}/{transport} done", file=sys.stderr, flush=True)
            if transport == "p2p":
                res[f"p2p_used_{name}"] = getattr(s, "_nx", None) is not None
                res[f"seed_misses_{name}"] = getattr(s, "last_seed_misses", None)
        os.environ["CLDRD_DIST_P2P"] = "1"
        Dh, Ih = s.search_host(xq, k)          # host in, host out: slices land in one shared page-locked block
        res[f"host_shared_{name}"] = getattr(s, "_host", None) is not None
        # the arrays belong to the caller: four more searches (reversed queries) while the first result is still
        # referenced -- the rotation of result sets runs out and the scratch-and-copy path is taken -- must not touch it
        keep = [s.search_host(np.ascontiguousarray(xq[::-1]), k) for _ in range(4)] if name in ("small", "big") else []
        if name == "manyq":     # batches reported to the caller as they land, in order, with the final rows
            seen = []
            Dq, Iq = s.search_to_host(q, k, on_batch=lambda b0, nb, Db, Ib: seen.append((b0, nb, Db.copy(), Ib.copy())))
            if rank == 0:
                res["on_batch_ok"] = [(b0, nb) for b0, nb, _, _ in seen] == [(0, 8192), (8192, nq - 8192)] and all(
                    np.array_equal(Db, Dh[b0:b0 + nb]) and np.array_equal(Ib, Ih[b0:b0 + nb]) for b0, nb, Db, Ib in seen) and \
                    np.array_equal(Dq, Dh) and np.array_equal(Iq, Ih)
        if rank == 0:
            for Dr, Ir in keep:
                res[f"host_owned_{name}"] = res.get(f"host_owned_{name}", True) and bool(
                    np.array_equal(Dr[::-1], Dh) and np.array_equal(Ir[::-1], Ih))
            del keep
            full = torch.from_numpy(xb).to(dev)
            one = CD.ShardedSearcher.from_rows(full, 0, n, scan="f16", id_map=id_map)
            one.world, one.rank = 1, 0
            D1, I1 = one.search(q, k)
            for transport, (D, I) in out.items():
                res[f"bit_equal_{name}_{transport}"] = bool(torch.equal(D, D1) and torch.equal(I, I1))
            res[f"bit_equal_{name}_host"] = bool(np.array_equal(Dh, D1.cpu().numpy()) and np.array_equal(Ih, I1.cpu().numpy()))
            if name == "big":
                D, I = out["p2p"]
                D_ref, I_ref = O.search(xb, ids, xq, k)
                r = O.compare_topk(D.cpu().numpy(), I.cpu().numpy(), D_ref, I_ref, *O.search(xb, ids, xq, k + 16, dtype=np.float64))
                res["oracle_ok"] = bool(r["ok"])
            one.shard.close()
        s.close()
        dist.barrier()
    if rank == 0:
        with open(args.out, "w") as f:
            json.dump(res, f)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
