"""BASELINE.json configs 3, 4 and 5 on one GPU (timing + checks; not the bench line).
  3: random-init DistilBERT (TAS-B shape) query encoding feeding top-1000 search, 6 980 queries
  4: bf16 index scan + fp32 rescore vs the fp32-stream (tf32) scan: overlap@1000
  5: curriculum data-gen shape: 502 939 queries x 8.8M passages, top-200 (streamed in 8192-query batches)
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cl-drd_b200")]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=8_841_823)
    ap.add_argument("--configs", default="3,4,5")
    ap.add_argument("--c5_queries", type=int, default=502_939)
    args = ap.parse_args()
    import torch
    from cldrd import dist as CD
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(1000)
    rows = torch.empty((args.rows, 768), dtype=torch.float32, device=dev)
    for r0 in range(0, args.rows, 1 << 20):
        rows[r0:r0 + (1 << 20)].normal_(generator=g)
    out = []
    todo = [int(c) for c in args.configs.split(",")]
    if 3 in todo:
        from transformers import DistilBertConfig, DistilBertModel
        torch.manual_seed(2)
        enc = DistilBertModel(DistilBertConfig()).to(dev).eval()          # 6 layers, 768, 12 heads, 66M params
        gen = torch.Generator().manual_seed(3)
        lens = torch.randint(4, 29, (6980,), generator=gen)
        ids = torch.zeros((6980, 30), dtype=torch.long)
        mask = torch.zeros((6980, 30), dtype=torch.long)
        for i, L in enumerate(lens.tolist()):
            ids[i, 0], ids[i, L + 1] = 101, 102
            ids[i, 1:L + 1] = torch.randint(1000, 30522, (L,), generator=gen)
            mask[i, :L + 2] = 1
        s = CD.ShardedSearcher.from_rows(rows, 0, args.rows, scan="f16")

        def encode():
            embs = []
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
                for b0 in range(0, 6980, 512):
                    o = enc(input_ids=ids[b0:b0 + 512].to(dev), attention_mask=mask[b0:b0 + 512].to(dev))[0][:, 0, :]
                    embs.append(o.float())
            return torch.cat(embs).contiguous()

        for _ in range(2):
            q = encode()
            s.local.search_device(q, 1000, translate_ids=False)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        q = encode()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        D, I = s.local.search_device(q, 1000, translate_ids=False)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        out.append({"config": 3, "encode_s": t1 - t0, "search_s": t2 - t1, "queries_per_s_end_to_end": 6980 / (t2 - t0),
                    "stats": s.shard.stats(), "note": "embeddings stay on the device between encoder and search"})
        print(json.dumps(out[-1]), flush=True)
        s.shard.close()
        del s, enc
    if 4 in todo:
        q = torch.randn((6980, 768), generator=g, dtype=torch.float32, device=dev)
        res = {}
        for scan in ("bf16", "tf32"):
            s = CD.ShardedSearcher.from_rows(rows, 0, args.rows, scan=scan)
            for _ in range(2):
                s.local.search_device(q, 1000, translate_ids=False)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            D, I = s.local.search_device(q, 1000, translate_ids=False)
            torch.cuda.synchronize()
            res[scan] = (D, I, time.perf_counter() - t0, s.shard.stats())
            s.shard.close()
            del s
        same_ids = bool(torch.equal(res["bf16"][1], res["tf32"][1]))
        same_scores = bool(torch.equal(res["bf16"][0], res["tf32"][0]))
        inter = (res["bf16"][1].unsqueeze(2) == res["tf32"][1][:64].unsqueeze(1)[:, :, :]).any(2).float().mean().item() if False else None
        out.append({"config": 4, "overlap_at_1000": 1.0 if same_ids else "differs", "ids_bit_equal": same_ids,
                    "scores_bit_equal": same_scores, "bf16_search_s": res["bf16"][2], "tf32_search_s": res["tf32"][2],
                    "bf16_stats": res["bf16"][3], "tf32_stats": res["tf32"][3]})
        print(json.dumps(out[-1]), flush=True)
    if 5 in todo:
        nq = args.c5_queries
        s = CD.ShardedSearcher.from_rows(rows, 0, args.rows, scan="f16")
        q = torch.randn((nq, 768), generator=g, dtype=torch.float32, device=dev)
        s.local.search_device(q[:8192], 200, translate_ids=False)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        D, I = s.local.search_device(q, 200, translate_ids=False)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        flops = 2.0 * nq * args.rows * 768
        out.append({"config": 5, "queries": nq, "k": 200, "search_s": dt, "queries_per_s": nq / dt, "TFLOPs_whole_search": flops / dt / 1e12,
                    "stats": s.shard.stats(), "sorted": bool((D[:, 1:] <= D[:, :-1]).all().item())})
        print(json.dumps(out[-1]), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "configs_345.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
