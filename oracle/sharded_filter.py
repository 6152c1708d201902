"""numpy model of the sharded search's filter decisions.  TEST INFRASTRUCTURE ONLY.

Not a restatement of the reference (which has no sharded search that works: the `co.shard=True` branch of
retriever/retrieval_utils.py:165-182 raises NameError): this models what OUR kernels decide, so that the
exactness argument of DESIGN.md sections 4, 5 and 7 can be attacked on the CPU with adversarial scan errors:

  levels_from_samples   cl-drd_b200/csrc/node.cuh    levels_seed_kernel
  shard_candidates      scan filter (score >= seed) + select_merge_kernel's band cut max(v_k - band, seed)
  count_levels          node.cuh  count_levels_peers_kernel
  cut_from_counts       select.cuh  rescore_sort_kernel (the counted cut, computed from the summed count planes)
  verify                node.cuh  merge_keys_kernel's seed check (verify_seed_kernel on the NCCL transport)

The one property everything rests on: a scan score differs from the exact fp32 score by at most eps
(band = 2 * eps).  `run` plays the whole protocol on exact scores S and scan scores S_hat and reports, per
query, whether the true top-k survived into the re-scored set.
"""
from __future__ import annotations

import numpy as np


def levels_from_samples(topj: np.ndarray) -> np.ndarray:
    """topj [parts, nq, J] (each row best first, -inf padded) -> [nq, J]: the J best of the union, best first."""
    parts, nq, J = topj.shape
    allv = np.moveaxis(topj, 0, 1).reshape(nq, parts * J)
    return -np.sort(-allv, axis=1)[:, :J]


def shard_candidates(s_hat: np.ndarray, seed: float, band: float, k: int) -> np.ndarray:
    """Indices (into this shard's rows) the shard keeps for one query: scan score >= seed, and if at least k
    of them exist, also >= (k-th best scan score of the shard) - band."""
    idx = np.nonzero(s_hat >= seed)[0]
    if idx.shape[0] >= k:
        vk = np.partition(s_hat[idx], idx.shape[0] - k)[idx.shape[0] - k]
        idx = idx[s_hat[idx] >= max(vk - band, seed)]
    return idx


def count_levels(cand_scores: np.ndarray, levels_q: np.ndarray) -> np.ndarray:
    """counts[b] = #candidates with scan score >= levels_q[b] (levels descending)."""
    return (cand_scores[None, :] >= levels_q[:, None]).sum(axis=1).astype(np.int64)


def cut_from_counts(total_counts: np.ndarray, levels_q: np.ndarray, band: float, k: int) -> float:
    hit = np.nonzero(total_counts >= k)[0]
    return float(levels_q[hit[0]] - band) if hit.shape[0] else -np.inf


def verify(kth_exact: float, seed: float, band: float) -> bool:
    """True = the seed is proven harmless for this query (k-th merged exact score clears seed + eps)."""
    return seed == -np.inf or kth_exact >= seed + 0.5 * band


def run(S: np.ndarray, S_hat: np.ndarray, eps: float, k: int, shards: int, J: int = 32, sample_stride: int = 64,
        use_cut: bool = True, lost_counts: float = 0.0, rng=None):
    """S, S_hat [nq, N] exact and scan scores with |S_hat - S| <= eps.  Rows are split into `shards`
    contiguous ranges; every shard samples every `sample_stride`-th row.  Returns a dict of per-query arrays:
    ok (true top-k inside the re-scored union), verified, rescored (rows re-scored over all shards),
    collected (rows above the seed over all shards).  lost_counts: probability that a shard's counts for a
    query are dropped (models a shard whose query went to the fallback: counts are only lower bounds)."""
    nq, N = S.shape
    assert np.abs(S_hat - S).max() <= eps * (1 + 1e-6)
    band = 2.0 * eps
    bounds = [(r * N) // shards for r in range(shards + 1)]
    topj = np.full((shards, nq, J), -np.inf, dtype=S_hat.dtype)
    for r in range(shards):
        samp = S_hat[:, bounds[r]:bounds[r + 1]:sample_stride]
        jj = min(J, samp.shape[1])
        topj[r, :, :jj] = -np.sort(-samp, axis=1)[:, :jj]
    levels = levels_from_samples(topj)
    out = {key: np.zeros(nq, dtype=np.int64) for key in ("rescored", "collected")}
    out["ok"] = np.zeros(nq, dtype=bool)
    out["verified"] = np.zeros(nq, dtype=bool)
    for q in range(nq):
        seed = levels[q, J - 1]
        cands, counts = [], np.zeros(J, dtype=np.int64)
        for r in range(shards):
            sh = S_hat[q, bounds[r]:bounds[r + 1]]
            out["collected"][q] += int((sh >= seed).sum())
            idx = shard_candidates(sh, seed, band, k)
            cands.append(idx + bounds[r])
            if rng is None or rng.random() >= lost_counts:
                counts += count_levels(sh[idx], levels[q])
        cut = cut_from_counts(counts, levels[q], band, k) if use_cut else -np.inf
        kept = np.concatenate([c[S_hat[q, c] >= cut] for c in cands])
        out["rescored"][q] = kept.shape[0]
        true_top = np.argsort(-S[q], kind="stable")[:k]
        kth_true = S[q, true_top[-1]]
        # ties at the k-th exact score may be resolved either way; require every row strictly above it
        must = true_top[S[q, true_top] > kth_true]
        out["ok"][q] = np.isin(must, kept).all()
        exact_kept = np.sort(S[q, kept])[::-1]
        kth_merged = exact_kept[k - 1] if exact_kept.shape[0] >= k else -np.inf
        out["verified"][q] = verify(kth_merged, seed, band)
        if out["verified"][q]:
            # a verified query must return exactly the k best exact scores
            out["ok"][q] = out["ok"][q] and kth_merged == kth_true
    return out
