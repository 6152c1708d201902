"""A/B tuning inside ONE process / one box (power-capped clocks make cross-run numbers noisy):
builds the configs[1] index once, then for each (run_len, growth) setting re-creates the shard
and reports per-chunk scan time and TFLOP/s."""
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
    ap.add_argument("--settings", default="0:0:0:0,0:0:0:1,0:0:0:0,0:0:0:1")
    args = ap.parse_args()
    import torch
    from cldrd import dist as CD
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(1000)
    rows = torch.empty((args.rows, 768), dtype=torch.float32, device=dev)
    for r0 in range(0, args.rows, 1 << 20):
        rows[r0:r0 + (1 << 20)].normal_(generator=g)
    q = torch.randn((args.queries, 768), generator=g, dtype=torch.float32, device=dev)
    for setting in args.settings.split(","):
        rl, gr, ch, tc2, pipe = (setting.split(":") + ["0", "0", "0"])[:5]
        os.environ["CLDRD_TC2"] = tc2
        os.environ["CLDRD_PIPELINE"] = pipe
        os.environ["CLDRD_RUN_LEN"] = rl
        os.environ["CLDRD_GROWTH"] = gr
        os.environ["CLDRD_SEED_CHUNKS"] = ch
        s = CD.ShardedSearcher.from_rows(rows, 0, args.rows, scan=args.scan)
        s.shard.set_profiling(True)
        best = None
        for i in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            s.local.search_device(q, args.k, translate_ids=False)
            e1.record()
            torch.cuda.synchronize()
            tot = e0.elapsed_time(e1)
            ms, nl = s.shard.scan_time()
            if i >= 1 and (best is None or tot < best[0]):
                best = (tot, ms, s.shard.scan_launches(), s.shard.stats())
        tot, ms, launches, st = best
        print("   issuer waits:", s.shard.wait_cycles())
        per = " ".join(f"{r}:{t:.2f}ms({2 * args.queries * r * 768 / t / 1e9:.0f}TF)" for r, t in launches[:12])
        print(f"run_len={rl} growth={gr} seed_chunks={ch} tc2={tc2} pipe={pipe}: total {tot:.1f} ms scan {ms:.1f} ms fallback {st['fallback_queries']} "
              f"({2 * args.queries * args.rows * 768 / ms / 1e9:.0f} TF) surv/q {st['survivors'] / args.queries:.0f} | {per}",
              flush=True)
        s.shard.close()
        del s


if __name__ == "__main__":
    main()
