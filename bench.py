#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on BASELINE.json's configs.

metric   : queries/s for top-1000 over a synthetic 8 841 823 x 768 fp32 index (configs[1],
           MS MARCO-dev shape, 6 980 queries), at --gpus N B200s of one box.  Other workloads
           (--workload): "curriculum" = configs[4] (502 939 queries, top-200, streamed in 8192-query
           batches), "encoder" = configs[2] (random-init DistilBERT query encoder feeding the search,
           embeddings stay on the device), "bf16" = configs[3] (bf16 scan + fp32 rescore vs the
           fp32-stream scan: overlap@1000).
step     : one search of all queries over the whole index.  N>1: every rank holds a row shard and makes one
           asynchronous call per 8192-query batch; sample scores, candidate counts and the re-scored lists
           are stored by the kernels into the peers' HBM over NVLink, flag barriers between the ranks, the
           merging ranks store their slices into rank 0's HBM or into page-locked host memory
           (DESIGN.md section 7).  No NCCL call inside the step.
value    : device-resident throughput (queries already in HBM), CUDA events, max over ranks.
e2e      : the same through the reference-facing call with HOST buffers: `index.search(x, k)`
           with numpy in / numpy out at N=1 (cldrd_search_host); at N>1 ShardedSearcher.search_host:
           every rank uploads the queries from pinned memory, every rank's merge kernel stores its slice
           into one shared page-locked host block, rank 0 reads numpy arrays; copies inside the timed region.
roofline : scan kernel (tcgen05 tiles + fused filter): algorithmic FLOPs 2*Q*N_shard*d per pass
           divided by the summed device time of the scan launches (CUDA events on the launching
           stream, recorded inside libcldrd), against MEASURED_PEAKS.json's sustained bf16 peak;
           step_frac / e2e_frac: the same FLOPs over the whole step / the whole end-to-end call.
parity   : after the timed loops (a checker, outside every timed region) 64 queries of the step's own
           output are compared with a brute-force fp64 search of the same rows (torch matmul per shard,
           gathered and merged on rank 0) through oracle.compare_topk; at N>1 the device result and the
           host result must agree bit for bit.  A failed check makes the run exit non-zero.
--impl reference : the CPU restatement of the reference search (oracle/cpu_baseline.py; faiss
           itself is absent and un-pinned) on the host cores, bounded sample per step.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cl-drd_b200")]

N_ROWS, DIM, N_QUERIES, TOPK = 8_841_823, 768, 6980, 1000
C5_QUERIES, C5_TOPK = 502_939, 200
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


def workload_shape(args):
    """(queries, k, config label) of the selected workload."""
    if args.workload == "curriculum":
        nq = args.queries if args.queries else C5_QUERIES
        return nq, C5_TOPK, f"configs[4]: synthetic {args.rows}x{DIM} fp32 index, {nq} queries (8192-query batches), top-{C5_TOPK}"
    nq = args.queries if args.queries else N_QUERIES
    label = {"dev": "configs[1]", "encoder": "configs[2] (DistilBERT query encoder -> search)",
             "bf16": "configs[3] (bf16 scan + fp32 rescore)"}[args.workload]
    return nq, TOPK, f"{label}: synthetic {args.rows}x{DIM} fp32 index, {nq} queries, top-{TOPK}"


def run_reference(args):
    """Reference arm: CPU restatement of the reference search on the host cores, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import cpu_baseline as CB
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    nq_full, k, label = workload_shape(args)
    sample_rows = args.rows // 16
    # size the query sample so that one step is a few seconds of CPU work
    probe = CB.time_cpu_search(args.rows, DIM, k, 65536, 128)
    est_qps_sample = probe["value"] * (args.rows / sample_rows)
    nq = int(min(nq_full, max(128, (est_qps_sample * 4.0) // 128 * 128)))
    xb, xq = CB.make_sample(sample_rows, DIM, nq)
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        CB.search_torch_cpu(xb, xq, k)
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
    per_step = sum(times) / len(times)
    scale = args.rows / sample_rows
    qps = nq / (per_step * scale)
    sample = (f"{nq} of the {nq_full} queries x {sample_rows} rows (1/16 of the index rows) per step, top-{k}, "
              f"torch-CPU sgemm+topk with {torch.get_num_threads()} threads; q/s scaled linearly in rows to the full index")
    line = {
        "impl": "reference", "metric": METRIC, "value": qps, "unit": "queries/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": label, "sample": sample},
        "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": sample},
        "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference search = faiss IndexFlatIP.search (absent, un-pinned upstream): CPU port of its semantics",
    }
    print(json.dumps(line), flush=True)


def brute_force_shard(torch, rows, row0, q_sel, kk):
    """Checker: fp64 scores of q_sel [m, d] against this rank's rows (plain torch matmul, chunked), the kk best per
    query by (score descending, row ascending).  Returns (scores f64 [m, kk], global rows i64 [m, kk])."""
    m, n = q_sel.shape[0], rows.shape[0]
    qd = q_sel.double()
    best_s = torch.empty((m, 0), dtype=torch.float64, device=rows.device)
    best_r = torch.empty((m, 0), dtype=torch.int64, device=rows.device)
    chunk = 1 << 17
    for r0 in range(0, n, chunk):
        c = min(chunk, n - r0)
        s = qd @ rows[r0:r0 + c].double().T
        r = (torch.arange(r0, r0 + c, device=rows.device, dtype=torch.int64) + row0).expand(m, c)
        cat_s, cat_r = torch.cat([best_s, s], 1), torch.cat([best_r, r], 1)     # rows stay ascending left to right
        order = torch.sort(-cat_s, dim=1, stable=True).indices[:, :kk]          # stable: ties keep the lower row first
        best_s, best_r = torch.gather(cat_s, 1, order), torch.gather(cat_r, 1, order)
    if best_s.shape[1] < kk:
        pad = kk - best_s.shape[1]
        best_s = torch.cat([best_s, best_s.new_full((m, pad), -float("inf"))], 1)
        best_r = torch.cat([best_r, best_r.new_full((m, pad), -1)], 1)
    return best_s, best_r


def parity_check(torch, dist, rank, world, dev, rows, row0, q_dev, k, ids_np, D_step, I_step, n_check=64, margin=16):
    """Rank 0 returns the parity record of the step's own output (D_step, I_step: numpy on rank 0)."""
    import numpy as np
    nq = q_dev.shape[0]
    sel = np.unique(np.linspace(0, nq - 1, n_check).astype(np.int64))
    if world > 1:   # queries either side of the slice boundaries of the merge
        sl = (min(nq, 8192) + world - 1) // world
        edge = np.array([b for r in range(1, world) for b in (r * sl - 1, r * sl) if 0 <= b < nq], dtype=np.int64)
        sel = np.unique(np.concatenate([sel, edge]))
    q_sel = q_dev[torch.from_numpy(sel).to(dev)]
    kk = k + margin
    s_loc, r_loc = brute_force_shard(torch, rows, row0, q_sel, kk)
    if world > 1:
        if rank == 0:
            all_s = [torch.empty_like(s_loc) for _ in range(world)]
            all_r = [torch.empty_like(r_loc) for _ in range(world)]
            dist.gather(s_loc, all_s, dst=0)
            dist.gather(r_loc, all_r, dst=0)
            cat_s, cat_r = torch.cat(all_s, 1), torch.cat(all_r, 1)            # shards in row order: rows ascending
            order = torch.sort(-cat_s, dim=1, stable=True).indices[:, :kk]
            s_loc, r_loc = torch.gather(cat_s, 1, order), torch.gather(cat_r, 1, order)
        else:
            dist.gather(s_loc, None, dst=0)
            dist.gather(r_loc, None, dst=0)
            return None
    from oracle import flat_ip as O     # the comparator only (tie bands, boundary allowance)
    s64 = s_loc.cpu().numpy()
    r64 = r_loc.cpu().numpy()
    ids_ext = np.where(r64 >= 0, ids_np[np.clip(r64, 0, None)], -1)
    D_ref = s64[:, :k].astype(np.float32)
    res = O.compare_topk(D_step[sel], I_step[sel], D_ref, ids_ext[:, :k], s64, ids_ext)
    return {"ok": bool(res["ok"]), "queries": int(sel.shape[0]), "overlap": res["overlap"], "max_rel_err": res["max_rel_err"],
            "bad_scores": res["bad_scores"], "bad_ids": res["bad_ids"], "exact_pos_frac": res["exact_pos_frac"],
            "checker": "torch fp64 matmul over every shard's rows + stable sort, merged on rank 0; oracle.compare_topk "
                       "(scores 1e-5 relative, ids exact outside near-tie bands)"}


def writer_rate(np, D, I, nq, k):
    """Lines/s of the native run-file writer on this step's result (retrieve_top_passages.py:90-109)."""
    import cldrd
    qids = np.arange(nq, dtype=np.int64) * 7 + 1_000_000
    out = {}
    for where in ("/dev/shm", tempfile.gettempdir()):
        if not os.path.isdir(where):
            continue
        path = os.path.join(where, f"cldrd_bench_run_{os.getpid()}.tsv")
        try:
            t0 = time.perf_counter()
            cldrd.write_run_file(path, qids, I, D)
            dt = time.perf_counter() - t0
            out[where] = {"lines_per_s": nq * k / dt, "seconds": dt, "bytes": os.path.getsize(path)}
        finally:
            if os.path.exists(path):
                os.unlink(path)
    return out


def index_load_rate(torch, dist, np, rank, world, dev, barrier, max_over_ranks, n_rows_file=1 << 20):
    """File -> HBM rate of cldrd_shard_load_file (read_index + index_cpu_to_gpu of the reference,
    retriever/retrieve_top_passages.py:85-86): rank 0 writes an IxMp{IxFI} file of n_rows_file x 768 rows through the
    streaming writer, every rank then loads its row range of it.  A side measurement: when no directory has room for
    the file (or writing it fails) every rank learns so and the record says why."""
    from cldrd import dist as CD
    from cldrd._lib import check, lib, ptr
    need = n_rows_file * DIM * 4 + n_rows_file * 8 + (64 << 20)
    plan = [None, None]        # [path, reason it was skipped]
    if rank == 0:
        for where in ("/dev/shm", tempfile.gettempdir()):
            try:
                vfs = os.statvfs(where)
                if os.path.isdir(where) and vfs.f_bavail * vfs.f_frsize > need:
                    plan[0] = os.path.join(where, f"cldrd_bench_load_{os.getpid()}.index")
                    break
            except OSError:
                continue
        if plan[0] is None:
            plan[1] = "no directory with room for the test file"
        else:
            try:
                w = C.c_void_p()
                check(lib().cldrd_index_writer_begin(C.byref(w), plan[0].encode(), n_rows_file, DIM, 1, 0))
                block = np.random.Generator(np.random.PCG64(5)).standard_normal((1 << 15, DIM), dtype=np.float32)
                for r0 in range(0, n_rows_file, 1 << 15):
                    check(lib().cldrd_index_writer_append(w, ptr(block), min(1 << 15, n_rows_file - r0)))
                ids = np.arange(n_rows_file, dtype=np.int64)
                check(lib().cldrd_index_writer_finish(w, ptr(ids)))
            except Exception as e:
                plan = [None, f"writing the test file failed: {e}"[:200]]
    if world > 1:
        dist.broadcast_object_list(plan, src=0)
    path = plan[0]
    if path is None:
        return {"skipped": plan[1]}
    try:
        barrier()
        t0 = time.perf_counter()
        s = CD.ShardedSearcher.from_file(path, dev.index, scan="f16")
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - t0)
        nbytes = n_rows_file * DIM * 4
        s.shard.close()
        barrier()
        return {"file_bytes": nbytes, "seconds": dt, "gb_per_s": nbytes / dt / 1e9, "ranks": world,
                "note": f"{os.path.dirname(path)} (page cache) -> pread -> pinned ring -> HBM, incl. the fp16 scan copy and the norm pass"}
    finally:
        if rank == 0 and os.path.exists(path):
            os.unlink(path)


def inprocess_e2e(torch, np, cldrd, CD, world, n_rows, d, scan, ids_np, q_np, k, steps, D_ref, I_ref):
    """`GpuIndexShards.search(numpy)` over all GPUs from this one process: queries/s end to end, and whether the result
    equals the one-process-per-GPU result bit for bit."""
    from cldrd.index import GpuIndexShards, shard_ranges
    parts = []
    for r, rr in enumerate(shard_ranges(n_rows, world)):
        dv = torch.device("cuda", r)
        with torch.cuda.device(dv):
            g = torch.Generator(device=dv).manual_seed(1000 + r)
            rows = torch.empty((len(rr), d), dtype=torch.float32, device=dv)
            for r0 in range(0, len(rr), 1 << 20):
                rows[r0:r0 + (1 << 20)].normal_(generator=g)
            parts.append(CD.ShardedSearcher.from_rows(rows, rr.start, n_rows, scan=scan))
    multi = GpuIndexShards([p.shard for p in parts], ids_np, n_rows, d)
    try:
        for _ in range(2):
            multi.search(q_np, k)
        D = I = None
        t0 = time.perf_counter()
        for _ in range(steps):
            D = I = None
            D, I = multi.search(q_np, k)
        dt = (time.perf_counter() - t0) / steps
        return {"value": q_np.shape[0] / dt, "unit": "queries/s", "ms_per_step": dt * 1e3,
                "equals_per_process_result_bitwise": bool(np.array_equal(D, D_ref) and np.array_equal(I, I_ref)),
                "api": "index_cpu_to_gpu_multiple(shard=True) -> GpuIndexShards.search(numpy): one process, one host thread, "
                       "all GPUs; merge kernels store straight into the caller's page-locked arrays"}
    finally:
        multi.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="dev", choices=["dev", "curriculum", "encoder", "bf16"])
    ap.add_argument("--scan", default=os.environ.get("CLDRD_BENCH_SCAN", "f16"))
    ap.add_argument("--rows", type=int, default=N_ROWS, help="(testing) override the index rows")
    ap.add_argument("--queries", type=int, default=0, help="(testing) override the query count")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip writer / index-load / batch-128 side measurements")
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
    n_rows, d = args.rows, DIM
    nq, k, label = workload_shape(args)
    warmup = max(args.warmup, 3)
    scan = "bf16" if args.workload == "bf16" else args.scan

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

    # ---- synthetic inputs of the named shape: N(0,1) rows generated per shard on the GPU --------
    rr = shard_ranges(n_rows, world)[rank]
    g = torch.Generator(device=dev).manual_seed(1000 + rank)
    rows = torch.empty((len(rr), d), dtype=torch.float32, device=dev)
    for r0 in range(0, len(rr), 1 << 20):
        rows[r0:r0 + (1 << 20)].normal_(generator=g)
    encode = None
    if args.workload == "encoder":
        # configs[2]: token rows [CLS] 4..28 random ids [SEP], pad 0, max_len 30; random-init DistilBertConfig(), seed 2,
        # autocast fp16, CLS pooling (models/nway_dual_encoder.py:51-57); the embeddings never leave the device
        from transformers import DistilBertConfig, DistilBertModel
        torch.manual_seed(2)
        enc = DistilBertModel(DistilBertConfig()).to(dev).eval()
        gen = torch.Generator().manual_seed(3)
        lens = torch.randint(4, 29, (nq,), generator=gen)
        tok = torch.zeros((nq, 30), dtype=torch.long)
        mask = torch.zeros((nq, 30), dtype=torch.long)
        for i, L in enumerate(lens.tolist()):
            tok[i, 0], tok[i, L + 1] = 101, 102
            tok[i, 1:L + 1] = torch.randint(1000, 30522, (L,), generator=gen)
            mask[i, :L + 2] = 1
        tok, mask = tok.pin_memory(), mask.pin_memory()

        def encode():
            embs = []
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
                for b0 in range(0, nq, 512):
                    o = enc(input_ids=tok[b0:b0 + 512].to(dev, non_blocking=True),
                            attention_mask=mask[b0:b0 + 512].to(dev, non_blocking=True))[0][:, 0, :]
                    embs.append(o.float())
            return torch.cat(embs).contiguous()

        q_dev = encode()
        q_host = q_dev.cpu().pin_memory()
    else:
        q_host = torch.randn((nq, d), generator=torch.Generator().manual_seed(1), dtype=torch.float32).pin_memory()
        q_dev = q_host.to(dev)
    ids_np = np.random.Generator(np.random.PCG64(7)).permutation(n_rows).astype(np.int64)
    id_map = torch.from_numpy(ids_np).to(dev) if rank == 0 else None
    searcher = CD.ShardedSearcher.from_rows(rows, rr.start, n_rows, scan=scan, id_map=id_map)
    shard = searcher.shard
    if world == 1:
        from cldrd._lib import check, lib, ptr
        check(lib().cldrd_shard_set_ids(shard.handle, ptr(ids_np)))  # N=1: ids applied inside the search
    torch.cuda.synchronize()

    # ---- device-resident loop: `value` ------------------------------------------------------------
    def step_dev():
        q = encode() if encode is not None else q_dev
        if world == 1:
            return searcher.local.search_device(q, k, translate_ids=True)
        return searcher.search(q, k)

    shard.set_profiling(True)
    for _ in range(warmup):
        step_dev()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()      # before the barrier: spawning nvidia-smi must not delay rank 0 inside the timed region
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    scan_ms_total, scan_launches, launches = 0.0, 0, 0
    D_dev = I_dev = None
    phase_hist = []
    e0.record()
    for _ in range(args.steps):
        D_dev = I_dev = None               # a caller's loop drops the previous result before it asks for the next one
        D_dev, I_dev = step_dev()
        if world > 1 and searcher.last_phase_ms:
            phase_hist.append(dict(searcher.last_phase_ms))
        ms, nl = shard.scan_time()
        scan_ms_total += ms
        scan_launches += nl
        launches += shard.stats()["launches"]
    e1.record()
    barrier()
    clocks = sampler.stop() if sampler else None
    shard.set_profiling(False)
    dev_ms = max_over_ranks(e0.elapsed_time(e1))
    stats = shard.stats()
    qps = nq * args.steps / (dev_ms / 1e3)
    phase_ms = getattr(searcher, "last_phase_ms", None)
    phase_by_rank = None
    if world > 1:
        # mean over the timed steps (last batch of each step), one record per rank: shows which GPU the others wait for
        mean = {k_: sum(p[k_] for p in phase_hist) / len(phase_hist) for k_ in phase_hist[0]} if phase_hist else phase_ms
        phase_by_rank = [None] * world
        dist.all_gather_object(phase_by_rank, mean)

    # ---- end-to-end loop with host buffers: `e2e` ---------------------------------------------------
    q_np = q_host.numpy()

    def step_e2e():
        if encode is not None and world == 1:
            # configs[2]: token ids from pinned host memory -> encoder -> search on the device-resident embeddings
            # (GpuIndexFlat.search takes the CUDA tensor: SURVEY f-3) -> numpy results
            return searcher.local.search(encode(), k)
        if world == 1:
            return searcher.local.search(q_np, k)          # numpy in -> numpy out (cldrd_search_host)
        # every rank uploads its replica of the queries from pinned memory; every rank's merge kernel stores its slice
        # of the result into one shared page-locked block over its own PCIe link; rank 0 gets numpy arrays
        return searcher.search_host(q_host, k)

    for _ in range(2):
        step_e2e()
    barrier()
    D_host = I_host = None
    t0 = time.perf_counter()
    for _ in range(args.steps):
        D_host = I_host = None             # a caller's loop drops the previous result: its page-locked arrays are recycled
        D_host, I_host = step_e2e()
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_qps = nq * args.steps / e2e_s

    # ---- parity of the step's own output (checker; outside every timed region) ----------------------------
    D_np = D_dev.cpu().numpy() if rank == 0 else None
    I_np = I_dev.cpu().numpy() if rank == 0 else None
    parity = parity_check(torch, dist, rank, world, dev, rows, rr.start, q_dev, k, ids_np, D_np, I_np)
    if rank == 0:
        same = bool(np.array_equal(D_np, D_host) and np.array_equal(I_np, I_host))
        parity["host_equals_device_bitwise"] = same
        parity["sorted"] = bool((np.diff(D_np, axis=1) <= 0).all())
        parity["ok"] = bool(parity["ok"] and same and parity["sorted"])

    extras = {}
    if not args.no_extras and args.workload == "dev":
        # the reference's own call pattern: index_retrieve(index, q, top_k, batch=128), i.e. one host-buffer
        # search per 128 queries (retriever/retrieval_utils.py:141-147), without its .tolist() boxing
        if world == 1:
            for q0 in range(0, min(nq, 512), 128):
                searcher.local.search(q_np[q0:q0 + 128], k)
            t0 = time.perf_counter()
            for q0 in range(0, nq, 128):
                searcher.local.search(q_np[q0:q0 + 128], k)
            dt = time.perf_counter() - t0
            extras["e2e_reference_loop_batch128"] = {"value": nq / dt, "unit": "queries/s", "ms_total": dt * 1e3,
                                                     "calls": (nq + 127) // 128,
                                                     "note": "one index pass per 128 queries: HBM-bound"}
        if rank == 0:
            try:
                extras["writer"] = writer_rate(np, D_host, I_host, nq, k)
            except Exception as e:          # a side measurement must not take the line down
                extras["writer"] = {"error": f"{type(e).__name__}: {e}"[:300]}
        extras["index_load"] = index_load_rate(torch, dist, np, rank, world, dev, barrier, max_over_ranks)
        if world > 1:
            # The faiss-shaped multi-GPU call of the reference, retrieval_utils.py:165-182: ONE process drives all N GPUs
            # (index_cpu_to_gpu_multiple(..., shard=True) -> index.search(numpy)).  Rank 0 builds it next to the ranks'
            # own shards (same rows, same seeds) while the other ranks wait on a host-side barrier.
            cpu_group = dist.new_group(backend="gloo")
            if rank == 0:
                try:
                    extras["e2e_inprocess"] = inprocess_e2e(torch, np, cldrd, CD, world, n_rows, d, scan, ids_np, q_np, k, args.steps,
                                                            D_host, I_host)
                except Exception as e:      # a side measurement must not take the line down
                    extras["e2e_inprocess"] = {"error": f"{type(e).__name__}: {e}"[:300]}
            dist.barrier(group=cpu_group)

    if not args.no_extras and args.workload == "curriculum":
        # the curriculum step end to end: host queries in, run file out, the file growing batch by batch behind the
        # search (retriever/retrieve_top_passages.py:88-109 on the 502 939 training queries)
        where = tempfile.gettempdir()
        for cand in ("/dev/shm", tempfile.gettempdir()):       # ~37 bytes per line
            try:
                vfs = os.statvfs(cand)
                if vfs.f_bavail * vfs.f_frsize > nq * k * 40 + (256 << 20):
                    where = cand
                    break
            except OSError:
                continue
        path = os.path.join(where, f"cldrd_bench_c5_{os.getpid()}.tsv")
        qids = np.arange(nq, dtype=np.int64) * 3 + 7
        D_host = I_host = None
        barrier()
        t0 = time.perf_counter()
        stream = cldrd.RunFileStream(path) if rank == 0 else None
        if world == 1:
            chunk = 8 * 8192
            for c0 in range(0, nq, chunk):
                Dc, Ic = searcher.local.search(q_np[c0:c0 + chunk], k)
                stream.put(qids[c0:c0 + chunk], Ic, Dc)
        else:
            searcher.search_host(q_host, k, on_batch=(lambda b0, nb, Db, Ib: stream.put(qids[b0:b0 + nb], Ib, Db)) if rank == 0 else None)
        werr = None
        if rank == 0:
            try:
                stream.close()
            except Exception as e:          # a side measurement must not take the line down (disk full, ...)
                werr = f"{type(e).__name__}: {e}"[:300]
        dt = max_over_ranks(time.perf_counter() - t0)
        if rank == 0:
            extras["e2e_with_run_file"] = {"error": werr} if werr else {
                "value": nq / dt, "unit": "queries/s", "seconds": dt, "lines": nq * k, "bytes": os.path.getsize(path), "where": where,
                "note": "search_host + RunFileStream: batch i is formatted and written while batch i+1 is searched"}
            if os.path.exists(path):
                os.unlink(path)

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
    # the ideal step at N GPUs: the whole job's FLOPs spread over N tensor-core peaks
    flops_job = 2.0 * nq * n_rows * d
    ideal_ms = flops_job / world / (peaks["tflops"] * 1e12) * 1e3
    roofline = {"bound": "tensor", "achieved": achieved_tf, "peak": peaks["tflops"], "unit": "TFLOP/s",
                "frac": achieved_tf / peaks["tflops"], "traffic": traffic,
                "step_frac": ideal_ms / (dev_ms / args.steps), "e2e_frac": ideal_ms / (e2e_s * 1e3 / args.steps),
                "frac_note": "frac: scan launches only (the seed-sample launch is inside the time, not inside the FLOPs, at every N); "
                             "step_frac / e2e_frac: 2*Q*N*d / n_gpus / peak over the whole device step / end-to-end call",
                "traffic_note": "dram read+write bytes of the full-index scan launch, ncu --set full (profiles/r01_scan_traffic.json); "
                                "algorithmic bytes of that launch: %d" % (n_rows * d * (2 if shard.scan in ("f16", "bf16") else 4)),
                "kernel": f"scan_tc2_kernel<{shard.scan}> (cta_group::2; seed-sample launch + full-shard filter launch per batch)",
                "peak_source": f"MEASURED_PEAKS.json bf16 sustained ({peaks['source']})",
                # short launches of a short step (N = 8: 8.5 ms of scan per 10.6 ms step) run above the sustained clocks,
                # so `frac` can pass 1 there: the burst figure of the same file is the ceiling of a single launch
                "peak_burst": peaks["tflops_burst"], "frac_vs_burst": achieved_tf / peaks["tflops_burst"],
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

    # N=1: the counters cover the whole search; N>1: the last 8192-query batch
    stats_q = max(nq if world == 1 else (nq - 1) % 8192 + 1, 1)
    ok = True
    if rank == 0:
        ok = parity["ok"]
        transport = ("single shard" if world == 1 else
                     "kernels store sample scores / counts / re-scored lists into the peers' HBM over NVLink (CUDA IPC), flag "
                     "barriers, merge kernels store their slices into rank 0's HBM (value) or page-locked host memory (e2e)"
                     if searcher._nx is not None else "NCCL all-gather + all-to-all + slice-wise merge + NCCL gather")
        line = {
            "metric": METRIC if args.workload != "curriculum" else "queries/s, top-200 over 8.8Mx768 fp32 flat IP index (curriculum data-gen shape)",
            "value": qps, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
            "warmup": warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None,
            "dtype": {"f16": "f16 tensor scan + f32 rescore", "bf16": "bf16 tensor scan + f32 rescore",
                      "tf32": "tf32 tensor scan + f32 rescore", "simt": "f32"}[shard.scan],
            "data": "synthetic",
            "config": {"workload": label,
                       "scan": shard.scan, "results": "exact fp32 (proven filter band + fp32 rescore)",
                       "parallelism": f"index rows sharded over {world} GPU(s); " + transport,
                       "l2": f"inputs larger than L2: each pass streams {scan_bytes / 1e9:.1f} GB of index rows per GPU",
                       "chunks_per_pass": stats["chunks"], "rescored_per_query": stats["rescored"] / stats_q,
                       "survivors_per_query": stats["survivors"] / stats_q, "fallback_queries": stats["fallback_queries"]},
            "clocks": clocks,
            "e2e": {"value": e2e_qps, "unit": "queries/s", "h2d_bytes_per_step": nq * d * 4 if encode is None else nq * 30 * 8 * 2,
                    "d2h_bytes_per_step": nq * k * 12, "ms_per_step": e2e_s * 1e3 / args.steps,
                    "api": "GpuIndexFlat.search(numpy) -> cldrd_search_host" if world == 1 else
                           "pinned host -> ShardedSearcher.search_host -> merge kernels store into a shared page-locked host block"},
            "gpu_launches": launches,
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "parity": parity,
        }
        if phase_ms:
            line["phase_ms_last_batch"] = phase_ms
        if phase_by_rank:
            line["phase_ms_mean_by_rank"] = [{k: round(v, 3) for k, v in p.items()} if p else None for p in phase_by_rank]
        line.update(extras)
        print(json.dumps(line), flush=True)
    if world > 1:
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
        dist.broadcast(flag, src=0)
        ok = bool(flag.item())
        searcher.close()
        dist.barrier()
        dist.destroy_process_group()
    if not ok:
        sys.exit(3)


if __name__ == "__main__":
    main()
