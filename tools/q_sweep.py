"""Q-sweep on configs[1]'s index: queries per pass from 1 to 6980, device-resident.
Shows the HBM-bound -> tensor-bound crossover.  Prints one JSON line per Q."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cl-drd_b200")]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scan", default="f16")
    ap.add_argument("--rows", type=int, default=8_841_823)
    ap.add_argument("--k", type=int, default=1000)
    ap.add_argument("--qs", default="1,8,64,128,416,1024,6980")
    args = ap.parse_args()
    import torch
    from cldrd import dist as CD
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1400.0}
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(1000)
    rows = torch.empty((args.rows, 768), dtype=torch.float32, device=dev)
    for r0 in range(0, args.rows, 1 << 20):
        rows[r0:r0 + (1 << 20)].normal_(generator=g)
    s = CD.ShardedSearcher.from_rows(rows, 0, args.rows, scan=args.scan)
    s.shard.set_profiling(True)
    out = []
    for Q in [int(x) for x in args.qs.split(",")]:
        q = torch.randn((Q, 768), generator=g, dtype=torch.float32, device=dev)
        for _ in range(3):
            s.local.search_device(q, args.k, translate_ids=False)
        reps = 5
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        scan_ms = 0.0
        for _ in range(reps):
            s.local.search_device(q, args.k, translate_ids=False)
            scan_ms += s.shard.scan_time()[0]
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        scan_ms /= reps
        gbs = s.shard.scan_bytes / (scan_ms / 1e3) / 1e9
        tf = 2.0 * Q * args.rows * 768 / (scan_ms / 1e3) / 1e12
        line = {"Q": Q, "ms_per_search": ms, "queries_per_s": Q / (ms / 1e3), "scan_ms": scan_ms, "scan_GBps": gbs,
                "hbm_frac": gbs / peaks["hbm_gbs"], "scan_TFLOPs": tf, "tensor_frac": tf / peaks["bf16_tflops_sustained"],
                "bound": "hbm" if gbs / peaks["hbm_gbs"] > tf / peaks["bf16_tflops_sustained"] else "tensor",
                "stats": s.shard.stats()}
        print(json.dumps(line), flush=True)
        out.append(line)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"q_sweep_{args.scan}.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
