"""Curriculum group files from a ranked run.

The reference only ships the CONSUMER of these files: `NwayDataset.create_from_relT_most_semi_hard_file`
(dataset/nway_dataset.py:213-261) reads one JSON object per line with the keys
`qid`, `relT_pids`, `most_hard_pids`, `semi_hard_pids`, and `__getitem__` (dataset/nway_dataset.py:32-71)
asserts the list lengths its `label_mode` expects (5/10/20/30 "relT" passages, the other 30-n split over the
two negative groups).  The scripts that PRODUCE the files from the top-200 retrieval runs (scripts/unity/…,
retrieve_top_passages.py with --top_k 200) are absent upstream, so the slicing rule is not pinned by code:
here it is explicit and parameterised - rank windows over the (re-)ranked list of each query - and the
defaults follow the shape the label modes fix (n_rel + n_most + n_semi = 30).

Everything works on the arrays `index_retrieve` returns, so a search result can be turned into a
training file without the run file in between (`groups_from_search`), or from a run file on disk
(`read_run` -> `build_groups`).
"""
from __future__ import annotations

import json
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

# (n_rel, n_most, n_semi) each `label_mode` of dataset/nway_dataset.py:41-70 accepts through
# create_from_relT_most_semi_hard_file (neg_pids = most_hard_pids + semi_hard_pids)
LABEL_MODE_SHAPES = {
    "2": (10, 10, 10), "3": (10, 10, 10), "4": (10, 10, 10), "9": (10, 10, 10),
    "5": (20, 5, 5), "10": (20, 5, 5),
    "6": (30, 0, 0),
    "7": (5, 12, 13), "8": (5, 12, 13),
}


def read_run_arrays(path, threads: int = 0) -> Tuple[np.ndarray, np.ndarray]:
    """(qid, pid) of every line of a run file, in file order, parsed by libcldrd on all host cores (cldrd_read_run:
    the reference's acceptance rule, 2 to 4 tab-separated fields after strip(), integer ids).  100 M lines (config 5's
    top-200 run of the 502 939 training queries) take seconds instead of the minutes of a Python line loop."""
    import ctypes as C
    from ._lib import E_FORMAT, CldrdError, check, lib, ptr
    n, bad = C.c_int64(), C.c_int64(-1)
    check(lib().cldrd_read_run(str(path).encode(), None, None, 0, int(threads), C.byref(n), None))
    q = np.empty((n.value,), dtype=np.int64)
    p = np.empty((n.value,), dtype=np.int64)
    n2 = C.c_int64()
    try:
        check(lib().cldrd_read_run(str(path).encode(), ptr(q), ptr(p), n.value, int(threads), C.byref(n2), C.byref(bad)))
    except CldrdError as e:
        if e.code == E_FORMAT:          # the reference's reader raises ValueError on such a line (retrieval_evaluator.py:55)
            raise ValueError(str(e)) from None
        raise
    if n2.value != n.value:
        raise RuntimeError(f"{path} changed while it was read ({n.value} -> {n2.value} lines)")
    return q, p


def read_run(path, threads: int = 0) -> Tuple[np.ndarray, List[np.ndarray]]:
    """Run file ("qid\\tpid[\\trank[\\tscore]]", what retrieve_top_passages.py:102-107 writes and
    evaluation/retrieval_evaluator.py:46-63 reads) -> (qids in first-seen order, pids per query in file order).
    A qid that comes back later in the file continues its list, like the `+=` of the reference's reader."""
    q, p = read_run_arrays(path, threads)
    if q.shape[0] == 0:
        return q, []
    starts = np.flatnonzero(np.concatenate(([True], q[1:] != q[:-1])))          # runs of equal qids
    heads = q[starts]
    if np.unique(heads).shape[0] == heads.shape[0]:                                 # the usual file: one run per query
        return heads, _cut(p, starts)
    # some qid has several runs: group them behind its first one, file order inside a group
    uniq, first, inv = np.unique(q, return_index=True, return_inverse=True)
    by_first = np.argsort(first, kind="stable")
    slot_of = np.empty_like(by_first)
    slot_of[by_first] = np.arange(by_first.shape[0])
    slot = slot_of[inv]
    order = np.argsort(slot, kind="stable")
    counts = np.bincount(slot, minlength=uniq.shape[0])
    return uniq[by_first], _cut(p[order], np.concatenate(([0], np.cumsum(counts)[:-1])))


def _cut(p: np.ndarray, starts: np.ndarray) -> List[np.ndarray]:
    """Views of p from every start to the next (np.split spends microseconds per piece; config 5 has 502 939)."""
    b = starts.tolist()
    e = b[1:] + [p.shape[0]]
    return [p[i:j] for i, j in zip(b, e)]


def rerank_with_teacher(qids: Sequence[int], ranked_pids: Sequence[np.ndarray], teacher_qids: Sequence[int],
                        teacher_ranked: Sequence[np.ndarray], keep_unscored: bool = True) -> List[np.ndarray]:
    """Teacher re-rank of the student's candidates before they are cut into groups (the "relT" lists of
    dataset/nway_dataset.py:239-261 are ranked by the TEACHER).  The teacher's order comes as a run file, which is
    what the reference's re-ranker writes: evaluation/reranking_evaluator.py:70-86 scores (query, passage) pairs and
    evaluation/utils.py:145-159 (`write_rankdata`) sorts them by score and writes "qid\tpid\trank\tscore".
    Per query: the student's candidates in teacher order; candidates the teacher did not score keep their student
    order behind the scored ones (keep_unscored) or are dropped; passages only the teacher lists are ignored
    (they were not retrieved).  A query the teacher file lacks keeps its student order."""
    t_of = {int(q): np.asarray(l, dtype=np.int64) for q, l in zip(teacher_qids, teacher_ranked)}
    out = []
    for qid, ranked in zip(qids, ranked_pids):
        ranked = np.asarray(ranked, dtype=np.int64)
        t = t_of.get(int(qid))
        if t is None:
            out.append(ranked)
            continue
        cand = set(ranked[ranked >= 0].tolist())
        scored = [p for p in t.tolist() if p in cand]
        seen = set(scored)
        rest = [p for p in ranked.tolist() if p >= 0 and p not in seen] if keep_unscored else []
        out.append(np.asarray(scored + rest, dtype=np.int64))
    return out


def _window(ranked: np.ndarray, lo: int, hi: int) -> np.ndarray:
    return ranked[min(lo, ranked.shape[0]):min(hi, ranked.shape[0])]


def _draw(pool: np.ndarray, keys: np.ndarray, n: int) -> np.ndarray:
    """n members of pool without replacement, in pool (= rank) order: those with the n smallest keys."""
    if n == 0:
        return pool[:0]
    m = pool.shape[0]
    pick = np.argpartition(keys[:m], n - 1)[:n] if n < m else np.arange(m)
    return pool[np.sort(pick)]


GROUP_BLOCK = 8192      # queries per block of random keys (part of what `seed` means: do not change)


def build_groups(qids: Sequence[int], ranked_pids: Sequence[np.ndarray], n_rel: int = 10, n_most: int = 10,
                 n_semi: int = 10, most_window: Optional[Tuple[int, int]] = None,
                 semi_window: Optional[Tuple[int, int]] = None, qrels: Optional[Dict[int, Iterable[int]]] = None,
                 seed: int = 0, strict: bool = True) -> List[dict]:
    """One example per query: the first `n_rel` ranks are the teacher-relevant group, `n_most` passages are
    drawn without replacement from ranks `most_window` (default: the n_rel..50 band) and `n_semi` from
    `semi_window` (default: 50..200).  Draws keep rank order.  Padding ids (-1, an index with fewer rows than
    top_k) are dropped first.  `qrels` (qid -> judged-relevant pids), when given, are moved to the front of the
    relevant group - the human positive is always a member, like `rel_pid` in the reference's other loaders
    (dataset/nway_dataset.py:204-209) - and never sampled as a negative.
    strict: a query whose lists cannot be filled raises (the dataset would assert later); else it is skipped.

    The draw: every query gets one uniform key per window position (PCG64(seed), blocks of GROUP_BLOCK queries) and
    takes the positions with the smallest keys.  Queries whose list needs no clean-up (no padding, no repeated pid, no
    judged positive, the block's common length) are cut as one matrix per block; the others one by one, by the same
    rule and with the same keys, so the result does not depend on which way a query went (config 5: 502 939 queries)."""
    assert n_rel >= 1 and n_most >= 0 and n_semi >= 0
    most_window = most_window or (n_rel, max(50, n_rel + n_most))
    semi_window = semi_window or (most_window[1], max(200, most_window[1] + n_semi))
    assert most_window[0] >= n_rel and semi_window[0] >= most_window[1], "windows must not overlap"
    rng = np.random.Generator(np.random.PCG64(seed))
    wm, ws = most_window[1] - most_window[0], semi_window[1] - semi_window[0]
    qids = [int(q) for q in (qids.tolist() if isinstance(qids, np.ndarray) else qids)]
    out: List[Optional[dict]] = []

    def one(qid: int, ranked: np.ndarray, km: np.ndarray, ks: np.ndarray) -> Optional[dict]:
        ranked = np.asarray(ranked, dtype=np.int64)
        ranked = ranked[ranked >= 0]
        _, first = np.unique(ranked, return_index=True)          # a reranked list may repeat a pid
        ranked = ranked[np.sort(first)]
        pos = [int(p) for p in qrels.get(qid, ())] if qrels else []
        if pos:
            ranked = ranked[~np.isin(ranked, pos)]
            take = max(n_rel - len(pos), 0)
            rel = (pos + ranked[:take].tolist())[:n_rel]
            # window positions stay "rank in the list with the positives moved to the front"
            ranked = np.concatenate([np.asarray(rel, dtype=np.int64), ranked[take:]])
        else:
            rel = ranked[:n_rel].tolist()
        most_pool = _window(ranked, *most_window)
        semi_pool = _window(ranked, *semi_window)
        if len(rel) < n_rel or most_pool.shape[0] < n_most or semi_pool.shape[0] < n_semi:
            if strict:
                raise ValueError(f"qid {qid}: {ranked.shape[0]} ranked passages cannot fill "
                                 f"{n_rel}+{n_most}+{n_semi} from windows {most_window}, {semi_window}")
            return None
        return {"qid": qid, "relT_pids": [int(p) for p in rel], "most_hard_pids": _draw(most_pool, km, n_most).tolist(),
                "semi_hard_pids": _draw(semi_pool, ks, n_semi).tolist()}

    for b0 in range(0, len(qids), GROUP_BLOCK):
        bq = qids[b0:b0 + GROUP_BLOCK]
        bl = [np.asarray(l, dtype=np.int64) for l in ranked_pids[b0:b0 + GROUP_BLOCK]]
        m = len(bq)
        km, ks = rng.random((m, wm)), rng.random((m, ws))
        res: List[Optional[dict]] = [None] * m
        done = np.zeros(m, dtype=bool)
        lens = np.fromiter((l.shape[0] for l in bl), dtype=np.int64, count=m)
        L = int(np.bincount(lens).argmax()) if m else 0
        plain = np.flatnonzero(lens == L)
        if qrels:
            plain = plain[[bq[i] not in qrels or not list(qrels[bq[i]]) for i in plain.tolist()]] if plain.size else plain
        if plain.size and L >= n_rel:
            R = np.stack([bl[i] for i in plain.tolist()])
            srt = np.sort(R, axis=1)
            ok = (R >= 0).all(axis=1) & ~(srt[:, 1:] == srt[:, :-1]).any(axis=1)
            plain, R = plain[ok], R[ok]
            mp, spl = R[:, most_window[0]:most_window[1]], R[:, semi_window[0]:semi_window[1]]
            if plain.size and mp.shape[1] >= n_most and spl.shape[1] >= n_semi:
                def draw(pool, keys, n):
                    if n == 0:
                        return pool[:, :0]
                    w = pool.shape[1]
                    pick = np.argpartition(keys[:, :w], n - 1, axis=1)[:, :n] if n < w else np.tile(np.arange(w), (pool.shape[0], 1))
                    return np.take_along_axis(pool, np.sort(pick, axis=1), axis=1)
                rel_l = R[:, :n_rel].tolist()
                most_l = draw(mp, km[plain], n_most).tolist()
                semi_l = draw(spl, ks[plain], n_semi).tolist()
                for j, i in enumerate(plain.tolist()):
                    res[i] = {"qid": bq[i], "relT_pids": rel_l[j], "most_hard_pids": most_l[j], "semi_hard_pids": semi_l[j]}
                done[plain] = True
        for i in np.flatnonzero(~done).tolist():
            res[i] = one(bq[i], bl[i], km[i], ks[i])
        out.extend(r for r in res if r is not None)
    return out


def groups_for_label_mode(qids, ranked_pids, label_mode: str, **kw) -> List[dict]:
    """Group sizes taken from the label mode the training run will use (dataset/nway_dataset.py:41-70)."""
    n_rel, n_most, n_semi = LABEL_MODE_SHAPES[str(label_mode)]
    return build_groups(qids, ranked_pids, n_rel, n_most, n_semi, **kw)


def groups_from_search(query_ids, nn_ids, **kw) -> List[dict]:
    """Straight from `index_retrieve`'s arrays (retriever/retrieval_utils.py:130-150): nn_ids [n, top_k]."""
    nn_ids = np.asarray(nn_ids, dtype=np.int64)
    return build_groups(np.asarray(query_ids, dtype=np.int64), list(nn_ids), **kw)


_GROUP_KEYS = ("qid", "relT_pids", "most_hard_pids", "semi_hard_pids")


def write_groups(path, examples: Iterable[dict]) -> int:
    """One JSON object per line, the layout dataset/nway_dataset.py:241-250 parses.  The text is json.dumps' (the
    repr of an int list IS its JSON text), written without a json.dumps call per example: config 5 has 502 939."""
    n = 0
    with open(path, "w") as f:
        for ex in examples:
            if tuple(ex) == _GROUP_KEYS and type(ex["qid"]) is int and all(
                    type(ex[k]) is list and all(type(v) is int for v in ex[k]) for k in _GROUP_KEYS[1:]):
                f.write(f'{{"qid": {ex["qid"]}, "relT_pids": {ex["relT_pids"]}, "most_hard_pids": {ex["most_hard_pids"]}, '
                        f'"semi_hard_pids": {ex["semi_hard_pids"]}}}\n')
            else:
                f.write(json.dumps(ex) + "\n")
            n += 1
    return n


def ranklists_for_evaluator(query_ids, nn_ids) -> Dict[int, List[int]]:
    """In-memory hand-off to `RankingEvaluator._calculate_metrics_plain` (evaluation/retrieval_evaluator.py:79):
    the `qid_to_ranklist` dict its `compute_metrics` builds from the run file (:46-63), without the file.
    A qid that occurs twice continues its list, as the reader's `+=` does; padding ids are dropped."""
    nn_ids = np.asarray(nn_ids, dtype=np.int64)
    out: Dict[int, List[int]] = {}
    for q, row in zip(np.asarray(query_ids, dtype=np.int64).tolist(), nn_ids):
        out.setdefault(q, []).extend(row[row >= 0].tolist())
    return out


def ranklists_from_run_file(path, threads: int = 0) -> Dict[int, List[int]]:
    """The `qid_to_ranklist` dict that `RankingEvaluator.compute_metrics` builds from a run file line by line
    (evaluation/retrieval_evaluator.py:46-63), built by the native reader instead: same keys in the same order, same
    lists (padding ids included, exactly what the reference's loop would hold).  Pass it to
    `_calculate_metrics_plain(ranklists, evaluator.qid_to_relevant_data, binarization_point=...)` (:64-76)."""
    qids, lists = read_run(path, threads)
    return {int(q): l.tolist() for q, l in zip(qids.tolist(), lists)}

