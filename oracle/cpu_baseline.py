"""CPU baseline for bench.py: the oracle's search restated with the fastest CPU primitives in the
image (torch CPU sgemm + topk, all host threads).  TEST/BENCH INFRASTRUCTURE ONLY.

The reference's own CPU search is faiss IndexFlatIP.search (retriever/retrieval_utils.py:135,143),
which cannot run here (faiss is absent and un-pinned upstream), so kind = "port".  Loop shape
follows index_retrieve(index, q, topk, batch=128) (retriever/retrieval_utils.py:131-153).
Same semantics as oracle/flat_ip.py (fp32 scores, descending, external ids); ties among equal
scores may come back in any order here, which is irrelevant for timing.
"""
from __future__ import annotations

import os
import time
from typing import Dict

import numpy as np
import torch


def make_sample(n_rows: int, d: int, nq: int, seed: int = 0):
    g = torch.Generator().manual_seed(seed)
    xb = torch.randn((n_rows, d), generator=g, dtype=torch.float32)
    xq = torch.randn((nq, d), generator=g, dtype=torch.float32)
    return xb, xq


def search_torch_cpu(xb: torch.Tensor, xq: torch.Tensor, k: int, batch: int = 128, row_block: int = 131072):
    """Blocked over rows so the score tile stays in cache-friendly size; per query batch the
    per-block top-k lists are merged, like faiss' CPU flat search does."""
    outs_D, outs_I = [], []
    for q0 in range(0, xq.shape[0], batch):
        q = xq[q0:q0 + batch]
        best_v = best_i = None
        for r0 in range(0, xb.shape[0], row_block):
            s = q @ xb[r0:r0 + row_block].T
            v, i = torch.topk(s, min(k, s.shape[1]), dim=1)
            i = i + r0
            if best_v is None:
                best_v, best_i = v, i
            else:
                cv, ci = torch.cat([best_v, v], 1), torch.cat([best_i, i], 1)
                best_v, sel = torch.topk(cv, min(k, cv.shape[1]), dim=1)
                best_i = torch.gather(ci, 1, sel)
        outs_D.append(best_v)
        outs_I.append(best_i)
    return torch.cat(outs_D), torch.cat(outs_I)


def time_cpu_search(n_rows_full: int, d: int, k: int, sample_rows: int, sample_queries: int,
                    repeats: int = 1) -> Dict:
    """Times `sample_queries` queries over `sample_rows` rows and scales linearly in rows to the
    full index (work per query is linear in N).  Returns queries/s over the FULL index."""
    torch.set_num_threads(os.cpu_count() or 1)
    xb, xq = make_sample(sample_rows, d, sample_queries)
    search_torch_cpu(xb[:65536], xq[:32], k)  # warm the thread pool
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        search_torch_cpu(xb, xq, k)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    scale = n_rows_full / sample_rows
    qps_full = sample_queries / (best * scale)
    return {
        "value": qps_full,
        "unit": "queries/s",
        "cores": torch.get_num_threads(),
        "kind": "port",
        "sample": f"{sample_queries} queries x {sample_rows} rows x {d} (1/{scale:.1f} of the index rows), "
                  f"top-{k}, torch-CPU sgemm+topk, {best:.2f} s, scaled linearly in rows",
        "seconds": best,
    }
