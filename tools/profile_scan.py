"""One full-size search for ncu: builds the configs[1] index on the GPU, warms up once, searches once.
Usage under ncu:  ncu ... python tools/profile_scan.py --scan f16 [--queries 6980]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cl-drd_b200")]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scan", default="f16")
    ap.add_argument("--rows", type=int, default=8_841_823)
    ap.add_argument("--queries", type=int, default=6980)
    ap.add_argument("--k", type=int, default=1000)
    ap.add_argument("--searches", type=int, default=2)
    args = ap.parse_args()
    import torch
    from cldrd import dist as CD
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(1000)
    rows = torch.empty((args.rows, 768), dtype=torch.float32, device=dev)
    for r0 in range(0, args.rows, 1 << 20):
        rows[r0:r0 + (1 << 20)].normal_(generator=g)
    q = torch.randn((args.queries, 768), generator=g, dtype=torch.float32, device=dev)
    s = CD.ShardedSearcher.from_rows(rows, 0, args.rows, scan=args.scan)
    s.shard.set_profiling(True)
    for i in range(args.searches):
        D, I = s.local.search_device(q, args.k, translate_ids=False)
        torch.cuda.synchronize()
        print("search", i, s.shard.stats(), "scan_ms", s.shard.scan_time())


if __name__ == "__main__":
    main()
