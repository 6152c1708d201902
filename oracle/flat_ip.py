"""CPU oracle for CL-DRD's dense-retrieval search path.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: the arithmetic of this path lives in `faiss` (facebookresearch/faiss,
`IndexIDMap(IndexFlatIP)`), a third-party dependency that the reference neither vendors nor
pins (no requirements file anywhere under /root/reference) and that is not installed in this
image.  The reference has no tests or golden vectors for the path either.
This file therefore restates the *published* semantics of faiss' flat inner-product index
and anchors on the reference's own call sites:

  * ``index.search(x, k)``           retriever/retrieval_utils.py:135,143
  * ``IndexIDMap(IndexFlatIP(d))``    retriever/index_text.py:91-97
  * ``write_index`` / ``read_index``  retriever/index_text.py:105, retriever/retrieve_top_passages.py:85
  * ``index_retrieve`` loop shape     retriever/retrieval_utils.py:131-153
  * regroup + run-file formatting     retriever/retrieve_top_passages.py:90-109
  * ``meta.pkl``                      retriever/index_text.py:107-109

What IS pinned: the restatements of the reference's own Python on this path -- ``index_retrieve``, ``regroup``,
``write_run``, ``write_meta`` -- reproduce, byte for byte, the files the reference's unmodified scripts wrote when they
were run in this container (tests/golden/make_golden_reference.py -> tests/golden/ref_pipeline/;
tests/test_oracle.py::test_oracle_restatements_reproduce_the_reference_run_here).  What stays unpinned: ``search`` and
the index-file bytes (faiss' own code), for which only the published semantics and the literal offsets exist.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import this module.  The product path (``cl-drd_b200/``) never does:
it has no CPU search at all and fails loudly when the CUDA library is missing.
"""
from __future__ import annotations

import io
import os
import pickle
import struct
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

NEG_FLT_MAX = np.float32(-3.4028234663852886e38)
METRIC_INNER_PRODUCT = 0
METRIC_L2 = 1

# --------------------------------------------------------------------------------------------
# a-6  index.search(x, k): S = X @ B.T in fp32, per-row top-k by descending score,
#      ties -> lower row, rows -> external ids through id_map, (-FLT_MAX, -1) padding if N < k.
#      (retriever/retrieval_utils.py:135,143; faiss IndexFlatIP / IndexIDMap semantics)
# --------------------------------------------------------------------------------------------


def _topk_desc_stable(scores: np.ndarray, k: int) -> Tuple[np.ndarray, np.ndarray]:
    """Per-row top-k of `scores` [n, N]: descending score, ties broken by lower column."""
    n, N = scores.shape
    kk = min(k, N)
    if kk < N:
        # partition first (cheap), then order the survivors with a stable sort on (-score, col)
        part = np.argpartition(-scores, kk - 1, axis=1)[:, :kk]
        # argpartition may split a tie group at the boundary arbitrarily: pull in every column
        # whose score equals the boundary score and redo the choice by lowest column.
        kth = np.take_along_axis(scores, part, axis=1).min(axis=1)
        rows_out = np.empty((n, kk), dtype=np.int64)
        for i in range(n):
            cand = np.nonzero(scores[i] >= kth[i])[0]  # ascending columns
            order = np.argsort(-scores[i, cand], kind="stable")[:kk]
            rows_out[i] = cand[order]
    else:
        rows_out = np.argsort(-scores, axis=1, kind="stable").astype(np.int64)
    vals = np.take_along_axis(scores, rows_out, axis=1)
    if kk < k:
        pad = k - kk
        vals = np.concatenate([vals, np.full((n, pad), NEG_FLT_MAX, dtype=scores.dtype)], axis=1)
        rows_out = np.concatenate([rows_out, np.full((n, pad), -1, dtype=np.int64)], axis=1)
    return vals, rows_out


def search_rows(xb: np.ndarray, xq: np.ndarray, k: int, dtype=np.float32,
                block: int = 65536) -> Tuple[np.ndarray, np.ndarray]:
    """Exhaustive inner-product top-k.  Returns (D [n,k] dtype, rows [n,k] int64, -1 padded).

    dtype=np.float32 is the reference behaviour (faiss sgemm); dtype=np.float64 is the
    "twin" used only to classify near-ties in the parity comparator.
    The index is walked in row blocks so that 8.8M-row slices fit in host memory; the block
    results are merged with the same (score desc, row asc) order, which makes the result
    independent of `block`.
    """
    xb = np.ascontiguousarray(xb)
    xq = np.ascontiguousarray(xq)
    assert xb.ndim == 2 and xq.ndim == 2 and xb.shape[1] == xq.shape[1]
    n, N = xq.shape[0], xb.shape[0]
    q = xq.astype(dtype, copy=False)
    best_v = np.empty((n, 0), dtype=dtype)
    best_r = np.empty((n, 0), dtype=np.int64)
    for r0 in range(0, max(N, 1), block):
        b = xb[r0:r0 + block].astype(dtype, copy=False)
        if b.shape[0] == 0:
            break
        s = q @ b.T
        v, r = _topk_desc_stable(s, min(k, b.shape[0]))
        r = r + r0
        best_v = np.concatenate([best_v, v], axis=1)
        best_r = np.concatenate([best_r, r], axis=1)
        if best_v.shape[1] > k:
            # rows inside best_r are ascending within equal scores only per block; a stable
            # sort on (-score) after ordering by row keeps "lower row first" globally.
            o = np.argsort(best_r, axis=1, kind="stable")
            best_v = np.take_along_axis(best_v, o, axis=1)
            best_r = np.take_along_axis(best_r, o, axis=1)
            o = np.argsort(-best_v, axis=1, kind="stable")[:, :k]
            best_v = np.take_along_axis(best_v, o, axis=1)
            best_r = np.take_along_axis(best_r, o, axis=1)
    if best_v.shape[1] < k:
        pad = k - best_v.shape[1]
        best_v = np.concatenate([best_v, np.full((n, pad), NEG_FLT_MAX, dtype=dtype)], axis=1)
        best_r = np.concatenate([best_r, np.full((n, pad), -1, dtype=np.int64)], axis=1)
    return best_v, best_r


def search(xb: np.ndarray, ids: Optional[np.ndarray], xq: np.ndarray, k: int,
           dtype=np.float32) -> Tuple[np.ndarray, np.ndarray]:
    """`IndexIDMap(IndexFlatIP).search`: rows translated through `ids` (label -1 kept)."""
    D, R = search_rows(xb, xq, k, dtype=dtype)
    if ids is None:
        return D, R
    ids = np.asarray(ids, dtype=np.int64)
    I = np.where(R >= 0, ids[np.clip(R, 0, None)], -1)
    return D, I


# --------------------------------------------------------------------------------------------
# a-5  index_retrieve(index, query_embeddings, topk, batch)   retriever/retrieval_utils.py:131-153
# --------------------------------------------------------------------------------------------


def index_retrieve(xb, ids, query_embeddings, topk, batch=None):
    """Loop/return shapes of the reference helper: batch=None -> ndarrays, else lists of lists.
    Return order is (scores, neighbours) as in the reference."""
    if batch is None:
        return search(xb, ids, query_embeddings, topk)
    nn_scores: List[List[float]] = []
    nearest: List[List[int]] = []
    off = 0
    while off < len(query_embeddings):
        qb = query_embeddings[off:off + batch]
        D, I = search(xb, ids, qb, topk)
        nearest.extend(I.tolist())
        nn_scores.extend(D.tolist())
        off += len(qb)
    return nn_scores, nearest


# --------------------------------------------------------------------------------------------
# a-2  index file: IndexIDMap{IndexFlatIP} as faiss.write_index lays it out (little-endian,
#      packed).  retriever/index_text.py:91-105, retriever/retrieve_top_passages.py:85.
# --------------------------------------------------------------------------------------------

_DUMMY = 1 << 20


def _index_header(fourcc: bytes, d: int, ntotal: int, metric: int) -> bytes:
    # fourcc, d i32, ntotal i64, dummy i64 x2, is_trained u8, metric_type i32
    return fourcc + struct.pack("<iqqqBi", d, ntotal, _DUMMY, _DUMMY, 1, metric)


def write_index_bytes(xb: np.ndarray, ids: Optional[np.ndarray], idmap2: bool = False) -> bytes:
    xb = np.ascontiguousarray(xb, dtype=np.float32)
    N, d = xb.shape
    out = io.BytesIO()
    if ids is not None:
        out.write(_index_header(b"IxM2" if idmap2 else b"IxMp", d, N, METRIC_INNER_PRODUCT))
    out.write(_index_header(b"IxFI", d, N, METRIC_INNER_PRODUCT))
    out.write(struct.pack("<Q", N * d))
    out.write(xb.tobytes())
    if ids is not None:
        ids = np.ascontiguousarray(ids, dtype=np.int64)
        assert ids.shape == (N,)
        out.write(struct.pack("<Q", N))
        out.write(ids.tobytes())
    return out.getvalue()


def write_index(path: str, xb: np.ndarray, ids: Optional[np.ndarray], idmap2: bool = False) -> None:
    with open(path, "wb") as f:
        f.write(write_index_bytes(xb, ids, idmap2))


def read_index_bytes(buf: bytes) -> Tuple[np.ndarray, Optional[np.ndarray], Dict]:
    off = 0

    def header(o):
        fourcc = buf[o:o + 4]
        d, ntotal, _d1, _d2, trained, metric = struct.unpack_from("<iqqqBi", buf, o + 4)
        o += 4 + 4 + 8 + 8 + 8 + 1 + 4
        if metric > 1:  # faiss writes an extra f32 metric_arg for the exotic metrics
            o += 4
        return fourcc, d, ntotal, metric, o

    fourcc, d, ntotal, metric, off = header(off)
    info = {"fourcc": fourcc.decode(), "d": d, "ntotal": ntotal, "metric": metric}
    has_map = fourcc in (b"IxMp", b"IxM2")
    if has_map:
        f2, d2, n2, m2, off = header(off)
        assert f2 in (b"IxFI", b"IxF2", b"IxFl"), f2
        assert d2 == d and n2 == ntotal
        info["inner_fourcc"] = f2.decode()
        info["metric"] = m2
    else:
        assert fourcc in (b"IxFI", b"IxF2", b"IxFl"), fourcc
    (cnt,) = struct.unpack_from("<Q", buf, off)
    off += 8
    assert cnt == ntotal * d
    info["data_off"] = off
    xb = np.frombuffer(buf, dtype="<f4", count=cnt, offset=off).reshape(ntotal, d).copy()
    off += 4 * cnt
    ids = None
    if has_map:
        (cnt2,) = struct.unpack_from("<Q", buf, off)
        off += 8
        assert cnt2 == ntotal
        info["ids_off"] = off
        ids = np.frombuffer(buf, dtype="<i8", count=cnt2, offset=off).copy()
        off += 8 * cnt2
    info["size"] = off
    return xb, ids, info


def read_index(path: str):
    with open(path, "rb") as f:
        return read_index_bytes(f.read())


def write_meta(index_dir: str, text_ids: Sequence[int]) -> None:
    """meta.pkl exactly as retriever/index_text.py:87,94,107-109 builds it."""
    text_ids_list = list(text_ids)
    text_id_to_idx = {tid: idx for idx, tid in enumerate(text_ids_list)}
    meta = {"text_ids": np.array(text_ids_list), "text_id_to_idx": text_id_to_idx}
    with open(os.path.join(index_dir, "meta.pkl"), "wb") as f:
        pickle.dump(meta, f)


# --------------------------------------------------------------------------------------------
# a-7/a-8  regroup + run file   retriever/retrieve_top_passages.py:90-109
# --------------------------------------------------------------------------------------------


def regroup(query_ids, nn_doc_ids, nn_scores):
    qid_to_ranks: Dict[int, list] = {}
    for qid, docids, scores in zip(query_ids, nn_doc_ids, nn_scores):
        for docid, s in zip(docids, scores):
            if qid not in qid_to_ranks:
                qid_to_ranks[qid] = [(docid, s)]
            else:
                qid_to_ranks[qid] += [(docid, s)]
    return qid_to_ranks


def write_run(path: str, query_ids, nn_doc_ids, nn_scores) -> float:
    """Pure-Python restatement of the reference writer (lists in, text out).  `nn_scores`
    must hold Python floats obtained by ``ndarray.tolist()`` like the reference does, so that
    the text is ``repr(float(np.float32))``."""
    if isinstance(nn_doc_ids, np.ndarray):
        nn_doc_ids = nn_doc_ids.tolist()
    if isinstance(nn_scores, np.ndarray):
        nn_scores = nn_scores.tolist()
    qid_to_ranks = regroup(list(query_ids), nn_doc_ids, nn_scores)
    total_rank = 0
    with open(path, "w") as f:
        for qid in qid_to_ranks:
            ranks = qid_to_ranks[qid]
            for i, (docid, s) in enumerate(ranks):
                f.write(f"{qid}\t{docid}\t{i+1}\t{s}\n")
            total_rank += len(ranks)
    return total_rank / max(len(qid_to_ranks), 1)


# --------------------------------------------------------------------------------------------
# Parity rule (north_star / SURVEY §8c): scores within 1e-5 relative; ids positionally equal
# except inside near-tie bands, which are compared as sets; overlap@k after that allowance.
# --------------------------------------------------------------------------------------------

REL_TOL = 1e-5


def compare_topk(D_test: np.ndarray, I_test: np.ndarray, D_ref: np.ndarray, I_ref: np.ndarray,
                 D_ref_ext: Optional[np.ndarray] = None, I_ref_ext: Optional[np.ndarray] = None,
                 rel_tol: float = REL_TOL) -> Dict[str, float]:
    """Tie-aware comparison.

    D_ref/I_ref: oracle top-k.  D_ref_ext/I_ref_ext (optional): oracle top-(k+margin) used to
    excuse swaps across the k-boundary when the boundary scores are inside the tolerance.
    Returns a dict of counters; `ok` is True when every difference is excused.
    """
    n, k = D_ref.shape
    assert D_test.shape == (n, k) and I_test.shape == (n, k)
    bad_scores = 0
    bad_ids = 0
    exact_pos = 0
    overlap_sum = 0.0
    max_rel = 0.0
    for i in range(n):
        dr, ir, dt, it = D_ref[i], I_ref[i], D_test[i], I_test[i]
        valid = ir >= 0
        # padding must agree exactly
        if not np.array_equal(it[~valid], ir[~valid]) or not np.array_equal(dt[~valid], dr[~valid]):
            bad_ids += int((~valid).sum())
        dr64 = dr[valid].astype(np.float64)
        dt64 = dt[valid].astype(np.float64)
        tol = rel_tol * np.maximum(np.abs(dr64), 1e-30)
        rel = np.abs(dt64 - dr64) / np.maximum(np.abs(dr64), 1e-30)
        if rel.size:
            max_rel = max(max_rel, float(rel.max()))
        bad_scores += int((np.abs(dt64 - dr64) > tol).sum())
        irv, itv = ir[valid], it[valid]
        exact_pos += int((irv == itv).sum())
        kk = irv.shape[0]
        if kk == 0:
            overlap_sum += 1.0
            continue
        # maximal runs whose adjacent reference scores differ by <= tol
        gaps = np.abs(np.diff(dr64))
        brk = np.nonzero(gaps > rel_tol * np.maximum(np.abs(dr64[:-1]), 1e-30))[0] + 1
        starts = np.concatenate([[0], brk])
        ends = np.concatenate([brk, [kk]])
        # Boundary band: ids whose fp64 score is chained (adjacent gaps <= tol) to the fp64
        # k-th score, on BOTH sides of the cut.  fp32 and fp64 may order such rows differently,
        # so either choice among them is a correct top-k.
        band: set = set()
        if D_ref_ext is not None and kk == k:
            de = D_ref_ext[i].astype(np.float64)
            ie = I_ref_ext[i]

            def close(a, b):
                return abs(a - b) <= rel_tol * max(abs(a), abs(b), 1e-30)

            j = k
            while j < de.shape[0] and ie[j] >= 0 and close(de[j], de[j - 1]):
                band.add(int(ie[j]))
                j += 1
            if j > k:
                band.add(int(ie[k - 1]))
                j = k - 1
                while j > 0 and close(de[j], de[j - 1]):
                    band.add(int(ie[j - 1]))
                    j -= 1
        miss = 0
        for s, e in zip(starts, ends):
            ref_set = set(irv[s:e].tolist())
            tst_set = set(itv[s:e].tolist())
            if ref_set == tst_set:
                continue
            if band and (ref_set ^ tst_set) <= band:
                continue
            miss += len(ref_set - tst_set)
        bad_ids += miss
        overlap_sum += 1.0 - miss / kk
    return {
        "ok": bad_scores == 0 and bad_ids == 0,
        "bad_scores": bad_scores,
        "bad_ids": bad_ids,
        "exact_pos_frac": exact_pos / max(n * k, 1),
        "overlap": overlap_sum / max(n, 1),
        "max_rel_err": max_rel,
    }


# --------------------------------------------------------------------------------------------
# Synthetic inputs of SURVEY §8(d): PCG64(seed).standard_normal(float32)
# --------------------------------------------------------------------------------------------


def synth(n: int, d: int, seed: int) -> np.ndarray:
    return np.random.Generator(np.random.PCG64(seed)).standard_normal((n, d), dtype=np.float32)


def synth_ids(n: int, seed: int = 7) -> np.ndarray:
    return np.random.Generator(np.random.PCG64(seed)).permutation(n).astype(np.int64)
