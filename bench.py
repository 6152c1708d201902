#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on BASELINE.json's config.

metric   : queries/s for top-1000 over a synthetic 8 841 823 x 768 fp32 index (configs[1],
           MS MARCO-dev shape, 6 980 queries), at --gpus N B200s of one box.
step     : one search of all 6 980 queries over the whole index.  N>1: every rank scans its row shard
           (seed and re-score cut agreed through an NCCL all-gather of sample scores and an all-reduce
           of 32 counts per query), the re-score kernel stores the lists into the merging rank's HBM
           over NVLink, slice-wise merge kernels, result on rank 0.
value    : device-resident throughput (queries already in HBM), CUDA events, max over ranks.
e2e      : the same through the reference-facing call with HOST buffers: `index.search(x, k)`
           with numpy in / numpy out at N=1 (cldrd_search_host); at N>1 ShardedSearcher.search_host:
           every rank uploads the queries from pinned memory and copies its merged slice into one
           shared page-locked host block, rank 0 reads numpy views; copies inside the timed region.
roofline : scan kernel (tcgen05 tiles + fused filter): algorithmic FLOPs 2*Q*N_shard*d per pass
           divided by the summed device time of the scan launches (CUDA events on the launching
           stream, recorded inside libcldrd), against MEASURED_PEAKS.json's sustained bf16 peak.
--impl reference : the CPU restatement of the reference search (oracle/cpu_baseline.py; faiss
           itself is absent and un-pinned) on the host cores, bounded sample per step.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cl-drd_b200")]

N_ROWS, DIM, N_QUERIES, TOPK = 8_841_823, 768, 6980, 1000
METRIC = "queries/s, top-1000 over 8.8Mx768 fp32 flat IP index"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "tflops": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "tflops_burst": d["bf16_tflops"], "source": "measured"}
    return {"hbm_gbs": 6650.0, "tflops": 1400.0, "tflops_burst": 1590.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (profiling recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(self.gpu)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, pw, reasons = [], [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
                pw.append(float(c[3]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm_sorted = sorted(sm)
        # median of the upper half = clocks under load (idle samples at the edges excluded)
        load = sm_sorted[len(sm_sorted) // 2:]
        return {"sm_mhz": load[len(load) // 2], "sm_max_mhz": max(mx), "power_w_max": max(pw),
                "samples": len(sm), "reasons": sorted(reasons)}


def run_reference(args):
    """Reference arm: CPU restatement of the reference search on the host cores, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cpu_baseline as CB
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    sample_rows = N_ROWS // 16
    # size the query sample so that one step is a few seconds of CPU work
    probe = CB.time_cpu_search(N_ROWS, DIM, TOPK, 65536, 128)
    est_qps_sample = probe["value"] * (N_ROWS / sample_rows)
    nq = int(min(N_QUERIES, max(128, (est_qps_sample * 4.0) // 128 * 128)))
    xb, xq = CB.make_sample(sample_rows, DIM, nq)
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        CB.search_torch_cpu(xb, xq, TOPK)
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
    per_step = sum(times) / len(times)
    scale = N_ROWS / sample_rows
    qps = nq / (per_step * scale)
    sample = (f"{nq} of the {N_QUERIES} queries x {sample_rows} rows (1/16 of the index rows) per step, top-{TOPK}, "
              f"torch-CPU sgemm+topk with {torch.get_num_threads()} threads; q/s scaled linearly in rows to the full index")
    line = {
        "impl": "reference", "metric": METRIC, "value": qps, "unit": "queries/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"configs[1]: synthetic {N_ROWS}x{DIM} fp32 index, {N_QUERIES} queries, top-{TOPK}",
                   "sample": sample},
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": sample},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference search = faiss IndexFlatIP.search (absent, un-pinned upstream): CPU port of its semantics",
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scan", default=os.environ.get("CLDRD_BENCH_SCAN", "f16"))
    ap.add_argument("--rows", type=int, default=N_ROWS, help="(testing) override the index rows")
    ap.add_argument("--queries", type=int, default=N_QUERIES, help="(testing) override the query count")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    import cldrd
    from cldrd import dist as CD
    from cldrd.index import shard_ranges

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device: there is no CPU path"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_rows, nq, k, d = args.rows, args.queries, TOPK, DIM
    warmup = max(args.warmup, 3)

    # ---- synthetic inputs of the named shape: N(0,1) rows generated per shard on the GPU --------
    rr = shard_ranges(n_rows, world)[rank]
    g = torch.Generator(device=dev).manual_seed(1000 + rank)
    rows = torch.empty((len(rr), d), dtype=torch.float32, device=dev)
    for r0 in range(0, len(rr), 1 << 20):
        rows[r0:r0 + (1 << 20)].normal_(generator=g)
    q_host = torch.randn((nq, d), generator=torch.Generator().manual_seed(1), dtype=torch.float32).pin_memory()
    q_dev = q_host.to(dev)
    ids_np = np.random.Generator(np.random.PCG64(7)).permutation(n_rows).astype(np.int64)
    id_map = torch.from_numpy(ids_np).to(dev) if rank == 0 else None
    searcher = CD.ShardedSearcher.from_rows(rows, rr.start, n_rows, scan=args.scan, id_map=id_map)
    shard = searcher.shard
    if world == 1:
        from cldrd._lib import check, lib, ptr
        check(lib().cldrd_shard_set_ids(shard.handle, ptr(ids_np)))  # N=1: ids applied inside the search
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- device-resident loop: `value` ------------------------------------------------------------
    def step_dev():
        if world == 1:
            return searcher.local.search_device(q_dev, k, translate_ids=True)
        return searcher.search(q_dev, k)

    for _ in range(warmup):
        step_dev()
    shard.set_profiling(True)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()      # before the barrier: spawning nvidia-smi must not delay rank 0 inside the timed region
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    scan_ms_total, scan_launches, launches = 0.0, 0, 0
    e0.record()
    for _ in range(args.steps):
        step_dev()
        ms, nl = shard.scan_time()
        scan_ms_total += ms
        scan_launches += nl
        launches += shard.stats()["launches"] + (2 if world > 1 else 0)   # + merge and verify of this rank's slice
    e1.record()
    barrier()
    clocks = sampler.stop() if sampler else None
    shard.set_profiling(False)
    dev_ms = max_over_ranks(e0.elapsed_time(e1))
    stats = shard.stats()
    qps = nq * args.steps / (dev_ms / 1e3)

    # ---- end-to-end loop with host buffers: `e2e` ---------------------------------------------------
    q_np = q_host.numpy()

    def step_e2e():
        if world == 1:
            return searcher.local.search(q_np, k)          # numpy in -> numpy out (cldrd_search_host)
        # every rank uploads its replica of the queries from pinned memory and writes its slice of the
        # result into one shared page-locked block over its own PCIe link; rank 0 gets numpy views
        return searcher.search_host(q_host, k)

    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_qps = nq * args.steps / e2e_s

    # the reference's own call pattern: index_retrieve(index, q, top_k, batch=128), i.e. one host-buffer
    # search per 128 queries (retriever/retrieval_utils.py:141-147), without its .tolist() boxing
    e2e_b128 = None
    if world == 1:
        for q0 in range(0, min(nq, 512), 128):
            searcher.local.search(q_np[q0:q0 + 128], k)
        t0 = time.perf_counter()
        for q0 in range(0, nq, 128):
            searcher.local.search(q_np[q0:q0 + 128], k)
        dt = time.perf_counter() - t0
        e2e_b128 = {"value": nq / dt, "unit": "queries/s", "ms_total": dt * 1e3, "calls": (nq + 127) // 128,
                    "note": "one index pass per 128 queries: HBM-bound"}

    # ---- roofline of the scan kernel ----------------------------------------------------------------
    peaks = load_peaks()
    flops_step_shard = 2.0 * nq * len(rr) * d
    scan_s = scan_ms_total / 1e3
    achieved_tf = flops_step_shard * args.steps / max(scan_s, 1e-9) / 1e12
    scan_bytes = shard.scan_bytes
    achieved_gbs = scan_bytes * args.steps / max(scan_s, 1e-9) / 1e9
    # worst rank bounds the job
    achieved_tf = -max_over_ranks(-achieved_tf)
    # DRAM traffic of the dominant launch from the committed ncu capture (same kernel, same shape);
    # only meaningful for the configuration that was captured: 1 GPU, f16 scan, full workload
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r01_scan_traffic.json")
    if world == 1 and shard.scan == "f16" and n_rows == N_ROWS and nq == N_QUERIES and os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f)["traffic_bytes_per_launch"]
    roofline = {"bound": "tensor", "achieved": achieved_tf, "peak": peaks["tflops"], "unit": "TFLOP/s",
                "frac": achieved_tf / peaks["tflops"], "traffic": traffic,
                "traffic_note": "dram read+write bytes of the full-index scan launch, ncu --set full (profiles/r01_scan_traffic.json); "
                                "algorithmic bytes of that launch: %d" % (n_rows * d * (2 if shard.scan in ("f16", "bf16") else 4)),
                "kernel": f"scan_tc2_kernel<{shard.scan}> (cta_group::2; seed-sample launch + full-index filter launch per step)",
                "peak_source": f"MEASURED_PEAKS.json bf16 sustained ({peaks['source']})",
                "flops_per_step_per_gpu": flops_step_shard, "scan_launches": scan_launches,
                "scan_ms_per_step": scan_ms_total / args.steps,
                "scan_share_of_step": (scan_ms_total / args.steps) / (dev_ms / args.steps),
                "hbm_gbs_algorithmic": achieved_gbs, "hbm_frac": achieved_gbs / peaks["hbm_gbs"]}

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import cpu_baseline as CB   # the one place bench may execute oracle/: the CPU baseline
        probe = CB.time_cpu_search(n_rows, d, k, 65536, 128)
        sample_rows = max(min(n_rows // 16, n_rows), 1)
        est = probe["value"] * (n_rows / sample_rows)
        sq = int(min(nq, max(128, (est * 10.0) // 128 * 128)))   # ~10 s of CPU work
        cpu_baseline = CB.time_cpu_search(n_rows, d, k, sample_rows, sq)
        cpu_baseline.pop("seconds", None)

    phase_ms = getattr(searcher, "last_phase_ms", None)
    if rank == 0:
        line = {
            "metric": METRIC, "value": qps, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
            "warmup": warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None,
            "dtype": {"f16": "f16 tensor scan + f32 rescore", "bf16": "bf16 tensor scan + f32 rescore",
                      "tf32": "tf32 tensor scan + f32 rescore", "simt": "f32"}[shard.scan],
            "data": "synthetic",
            "config": {"workload": f"configs[1]: synthetic {n_rows}x{d} fp32 index, {nq} queries, top-{k}",
                       "scan": shard.scan, "results": "exact fp32 (proven filter band + fp32 rescore)",
                       "parallelism": (f"index rows sharded over {world} GPU(s); " +
                                       ("single shard" if world == 1 else
                                        "re-score kernel stores lists into the merging rank's HBM over NVLink (peer memory), "
                                        "slice-wise merge kernels store into rank 0" if getattr(searcher, "_px", None) is not None
                                        else "NCCL all-to-all + slice-wise merge + NCCL gather")),
                       "l2": f"inputs larger than L2: each pass streams {scan_bytes / 1e9:.1f} GB of index rows per GPU",
                       "chunks_per_pass": stats["chunks"], "rescored_per_query": stats["rescored"] / max(nq, 1),
                       "survivors_per_query": stats["survivors"] / max(nq, 1), "fallback_queries": stats["fallback_queries"]},
            "clocks": clocks,
            "e2e": {"value": e2e_qps, "unit": "queries/s", "h2d_bytes_per_step": nq * d * 4,
                    "d2h_bytes_per_step": nq * k * 12, "ms_per_step": e2e_s * 1e3 / args.steps,
                    "api": "GpuIndexFlat.search(numpy) -> cldrd_search_host" if world == 1 else
                           "pinned host -> ShardedSearcher.search_host -> shared page-locked host block (numpy views on rank 0)"},
            "gpu_launches": launches,
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
        }
        if phase_ms:
            line["phase_ms_last_step"] = phase_ms
        if e2e_b128:
            line["e2e_reference_loop_batch128"] = e2e_b128
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
