"""Run-file writer: the regroup + write loops of retriever/retrieve_top_passages.py:90-109
(and retrieve_top_queries.py:65-82) as one native call, byte-identical output."""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

from ._lib import check, lib, ptr


def format_score(s: float) -> str:
    """Text the reference writes for one fp32 score: repr(float(np.float32(s)))."""
    buf = C.create_string_buffer(40)
    n = lib().cldrd_format_score(C.c_float(float(s)), buf)
    return buf.raw[:n].decode()


def write_run_file(path, query_ids, nn_ids, nn_scores, append: bool = False, threads: int = 0) -> float:
    """query_ids [n], nn_ids [n,k] int64, nn_scores [n,k] float32 -> "qid\\tdocid\\trank\\tscore\\n".

    Mirrors the reference: creates the parent directory if missing (:99-100); a query id that
    occurs more than once keeps its first position and its later hits continue the same rank
    sequence (the dict regroup of :90-96).  Returns the average ranks per query it prints (:109).
    threads: formatting threads (0 = CLDRD_WRITER_THREADS or every host core); the bytes do not depend on it.
    """
    qids = np.ascontiguousarray(np.asarray(query_ids, dtype=np.int64))
    I = np.ascontiguousarray(np.asarray(nn_ids, dtype=np.int64))
    D = np.ascontiguousarray(np.asarray(nn_scores, dtype=np.float32))
    assert I.ndim == 2 and D.shape == I.shape and qids.shape == (I.shape[0],)
    n, k = I.shape
    parent = Path(path).parent
    if not os.path.exists(parent):
        os.mkdir(parent)
    uniq, first = np.unique(qids, return_index=True)
    if uniq.shape[0] != n:
        # group duplicates behind their first occurrence, stable within a group
        first_of = dict(zip(uniq.tolist(), first.tolist()))
        order = np.argsort(np.array([first_of[q] for q in qids.tolist()], dtype=np.int64), kind="stable")
        qids, I, D = qids[order], np.ascontiguousarray(I[order]), np.ascontiguousarray(D[order])
    lines = C.c_int64()
    check(lib().cldrd_write_run_mt(str(path).encode(), ptr(qids), ptr(D), ptr(I), n, k, 1 if append else 0,
                                   int(threads), C.byref(lines)))
    return lines.value / max(uniq.shape[0], 1)


class RunFileStream:
    """Run file written while the search is still going: `put` hands one block of results (queries in file order) to
    a writer thread that appends it with the native multi-threaded formatter; the GIL is released inside the native
    call, so formatting and writing overlap the next search.  The bytes equal one write_run_file call over the
    concatenated blocks as long as no query id occurs in two blocks (the retriever's ids are unique: the dataset
    de-duplicates them, dataset/sequence_dataset.py:41-52).

    At config 5 (502 939 queries x 200 = 100 M lines) the reference's loop (retrieve_top_passages.py:90-109) would run
    for minutes after the search; here the file is complete a fraction of a second after the last batch lands."""

    def __init__(self, path, threads: int = 0):
        import queue
        import threading
        self.path = str(path)
        parent = Path(path).parent
        if not os.path.exists(parent):
            os.mkdir(parent)
        open(self.path, "wb").close()            # "w": truncate, like the reference
        self.threads = int(threads)
        self.lines = 0
        self.queries = 0
        self.error = None
        self._q = queue.Queue(maxsize=8)
        self._t = threading.Thread(target=self._work, daemon=True)
        self._t.start()

    def _work(self):
        while True:
            item = self._q.get()
            if item is None:
                return
            if self.error is not None:
                continue
            qids, I, D = item
            try:
                lines = C.c_int64()
                check(lib().cldrd_write_run_mt(self.path.encode(), ptr(qids), ptr(D), ptr(I), I.shape[0], I.shape[1], 1,
                                               self.threads, C.byref(lines)))
                self.lines += lines.value
                self.queries += I.shape[0]
            except Exception as e:      # surfaced by close()
                self.error = e

    def put(self, query_ids, nn_ids, nn_scores) -> None:
        qids = np.ascontiguousarray(np.asarray(query_ids, dtype=np.int64))
        I = np.ascontiguousarray(nn_ids, dtype=np.int64)
        D = np.ascontiguousarray(nn_scores, dtype=np.float32)
        assert I.ndim == 2 and D.shape == I.shape and qids.shape == (I.shape[0],)
        self._q.put((qids, I, D))

    def close(self) -> float:
        """Waits for the writer; returns the average ranks per query (the figure the reference prints, :109)."""
        self._q.put(None)
        self._t.join()
        if self.error is not None:
            raise self.error
        return self.lines / max(self.queries, 1)
