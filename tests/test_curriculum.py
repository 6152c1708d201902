"""curriculum group files and the in-memory evaluator hand-off (host logic, CPU)."""
import importlib.util
import json
import os
import sys
import types

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cl-drd_b200")]
REF = "/root/reference"

from cldrd import curriculum as CU  # noqa: E402


def _ranked(nq=12, depth=200, seed=3):
    rng = np.random.Generator(np.random.PCG64(seed))
    return np.arange(1000, 1000 + nq, dtype=np.int64), [rng.permutation(50_000)[:depth].astype(np.int64) for _ in range(nq)]


@pytest.mark.parametrize("mode", sorted(CU.LABEL_MODE_SHAPES, key=int))
def test_group_shapes_follow_label_mode(mode):
    qids, lists = _ranked()
    ex = CU.groups_for_label_mode(qids, lists, mode, seed=1)
    n_rel, n_most, n_semi = CU.LABEL_MODE_SHAPES[mode]
    assert len(ex) == len(qids)
    for e, ranked in zip(ex, lists):
        assert e["relT_pids"] == ranked[:n_rel].tolist()
        assert len(e["most_hard_pids"]) == n_most and len(e["semi_hard_pids"]) == n_semi
        pos = {int(p): i for i, p in enumerate(ranked)}
        most_r = [pos[p] for p in e["most_hard_pids"]]
        semi_r = [pos[p] for p in e["semi_hard_pids"]]
        assert most_r == sorted(most_r) and semi_r == sorted(semi_r)            # rank order kept
        assert all(n_rel <= r < 50 for r in most_r) and all(50 <= r < 200 for r in semi_r)
        allp = e["relT_pids"] + e["most_hard_pids"] + e["semi_hard_pids"]
        assert len(set(allp)) == len(allp) == 30


def test_deterministic_and_seeded():
    qids, lists = _ranked()
    a = CU.build_groups(qids, lists, seed=5)
    assert a == CU.build_groups(qids, lists, seed=5)
    assert a != CU.build_groups(qids, lists, seed=6)


def test_positives_lead_the_relevant_group_and_are_never_negatives():
    qids, lists = _ranked(nq=4)
    qrels = {int(qids[0]): [int(lists[0][120])],          # judged passage deep in the list
             int(qids[1]): [777_777],                     # judged passage not retrieved at all
             int(qids[2]): [int(lists[2][0])]}            # already on top
    ex = CU.build_groups(qids, lists, qrels=qrels, seed=0)
    assert ex[0]["relT_pids"][0] == int(lists[0][120]) and ex[0]["relT_pids"][1:] == lists[0][:9].tolist()
    assert ex[1]["relT_pids"][0] == 777_777
    assert ex[2]["relT_pids"] == lists[2][:10].tolist()
    assert ex[3]["relT_pids"] == lists[3][:10].tolist()
    for e, q in zip(ex, qids):
        for p in qrels.get(int(q), []):
            assert p not in e["most_hard_pids"] + e["semi_hard_pids"]


def test_padding_duplicates_and_short_lists():
    qids = np.array([1, 2], dtype=np.int64)
    full = np.arange(200, dtype=np.int64)
    short = np.concatenate([np.arange(40, dtype=np.int64), np.full(160, -1, dtype=np.int64)])   # index had 40 rows
    with pytest.raises(ValueError):
        CU.build_groups(qids, [full, short])
    ex = CU.build_groups(qids, [full, short], strict=False)
    assert [e["qid"] for e in ex] == [1]
    dup = np.concatenate([np.arange(100), np.arange(100), np.arange(100, 300)]).astype(np.int64)
    e = CU.build_groups([7], [dup])[0]
    allp = e["relT_pids"] + e["most_hard_pids"] + e["semi_hard_pids"]
    assert len(set(allp)) == 30 and e["relT_pids"] == list(range(10))


def test_run_file_round_trip(tmp_path):
    qids, lists = _ranked(nq=5, depth=200)
    run = tmp_path / "top200.run"
    with open(run, "w") as f:
        for q, l in zip(qids, lists):
            for r, p in enumerate(l):
                f.write(f"{q}\t{p}\t{r + 1}\t{1.0 / (r + 1)}\n")
    q2, l2 = CU.read_run(run)
    assert q2.tolist() == qids.tolist() and all((a == b).all() for a, b in zip(l2, lists))
    with open(tmp_path / "bad.run", "w") as f:
        f.write("1\n")
    with pytest.raises(ValueError):
        CU.read_run(tmp_path / "bad.run")


def test_cli_writes_what_the_loader_parses(tmp_path):
    spec = importlib.util.spec_from_file_location("mk", os.path.join(ROOT, "cl-drd_b200", "retriever", "make_curriculum_groups.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    qids, lists = _ranked(nq=6)
    run, qrels, out = tmp_path / "r.run", tmp_path / "qrels.tsv", tmp_path / "groups.json"
    with open(run, "w") as f:
        for q, l in zip(qids, lists):
            for r, p in enumerate(l):
                f.write(f"{q}\t{p}\t{r + 1}\t0.5\n")
    with open(qrels, "w") as f:
        f.write(f"{qids[0]}\t0\t{lists[0][33]}\t1\n{qids[1]}\t0\t{lists[1][3]}\t0\n")
    mk.main(mk.get_args(["--run_path", str(run), "--output_path", str(out), "--label_mode", "8", "--qrels_path", str(qrels)]))
    rows = [json.loads(l) for l in open(out)]
    assert len(rows) == 6 and rows[0]["relT_pids"][0] == int(lists[0][33])
    assert rows[1]["relT_pids"] == lists[1][:5].tolist()       # rel=0 in the qrels is not a positive
    assert all(len(r["relT_pids"]) == 5 and len(r["most_hard_pids"]) == 12 and len(r["semi_hard_pids"]) == 13 for r in rows)


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present on this box")
@pytest.mark.parametrize("mode", ["9", "8", "10", "6"])
def test_reference_nway_dataset_accepts_our_file(tmp_path, mode, monkeypatch):
    """dataset/nway_dataset.py:213-261 loads the file; __getitem__ (:32-71) asserts the group sizes of the label mode."""
    if "ujson" not in sys.modules:
        try:
            import ujson  # noqa: F401
        except ImportError:
            monkeypatch.setitem(sys.modules, "ujson", types.SimpleNamespace(loads=json.loads, dumps=json.dumps))
    pkg = types.ModuleType("refdataset")
    pkg.__path__ = [os.path.join(REF, "dataset")]
    monkeypatch.setitem(sys.modules, "refdataset", pkg)
    spec = importlib.util.spec_from_file_location("refdataset.nway_dataset", os.path.join(REF, "dataset", "nway_dataset.py"))
    mod = importlib.util.module_from_spec(spec)
    monkeypatch.setitem(sys.modules, "refdataset.nway_dataset", mod)
    spec.loader.exec_module(mod)

    qids, lists = _ranked(nq=5)
    out = tmp_path / "groups.json"
    CU.write_groups(out, CU.groups_for_label_mode(qids, lists, mode, seed=2))
    queries, passages = tmp_path / "q.tsv", tmp_path / "p.tsv"
    with open(queries, "w") as f:
        for q in qids:
            f.write(f"{q}\tquery {q}\n")
    with open(passages, "w") as f:
        for p in sorted({int(p) for l in lists for p in l}):
            f.write(f"{p}\tpassage {p}\n")
    ds = mod.NwayDataset.create_from_relT_most_semi_hard_file(str(queries), str(passages), str(out), tokenizer=None,
                                                              max_query_len=16, max_passage_len=32, label_mode=mode)
    assert len(ds) == 5
    n_rel, n_most, n_semi = CU.LABEL_MODE_SHAPES[mode]
    for i in range(len(ds)):
        item = ds[i]                                         # raises if the sizes do not fit the label mode
        assert len(item["relT_pids"]) == n_rel and len(item["neg_pids"]) == n_most + n_semi
        assert len(item["labels"]) == 30 and item["query"] == f"query {item['qid']}"


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present on this box")
def test_in_memory_hand_off_matches_run_file_metrics(tmp_path):
    """f-1: RankingEvaluator._calculate_metrics_plain (evaluation/retrieval_evaluator.py:79) on our dict ==
    compute_metrics on the run file."""
    sys.path.insert(0, ROOT)
    from oracle import flat_ip as O
    spec = importlib.util.spec_from_file_location("ref_eval2", os.path.join(REF, "evaluation", "retrieval_evaluator.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    xb, xq, ids = O.synth(3000, 32, 0), O.synth(25, 32, 1), O.synth_ids(3000)
    D, I = O.search(xb, ids, xq, 100)
    qids = np.arange(500, 525, dtype=np.int64)
    run = tmp_path / "dev.run"
    with open(run, "w") as f:
        for q, row in zip(qids, I):
            for r, p in enumerate(row):
                f.write(f"{q}\t{p}\t{r + 1}\n")
    qrels = tmp_path / "qrels.tsv"
    with open(qrels, "w") as f:
        for i, q in enumerate(qids):
            f.write(f"{q}\t0\t{I[i, (3 * i) % 11]}\t1\n")
    ev = mod.RankingEvaluator(str(qrels))
    from_file = ev.compute_metrics(str(run))
    ranklists = CU.ranklists_for_evaluator(qids, I)
    in_mem = ev._calculate_metrics_plain(ranklists, ev.qid_to_relevant_data, binarization_point=1.)
    in_mem = in_mem[0] if isinstance(in_mem, tuple) else in_mem
    for key, v in from_file.items():
        assert in_mem[key] == v, key
    # and the dict built from the run file by the native reader: the very dict compute_metrics builds for itself
    from_native = CU.ranklists_from_run_file(run)
    assert list(from_native) == qids.tolist() and from_native == {int(q): row.tolist() for q, row in zip(qids, I)}
    m2 = ev._calculate_metrics_plain(from_native, ev.qid_to_relevant_data, binarization_point=1.)
    m2 = m2[0] if isinstance(m2, tuple) else m2
    assert m2 == from_file


def _teacher_file(path, qids, lists, seed=11, drop_every=7):
    """A teacher's scores for the student's candidates, written the way evaluation/utils.py:145-159 (write_rankdata)
    writes them: per query sorted by score, "qid\\tpid\\trank\\tscore".  Every `drop_every`-th candidate is left unscored."""
    rng = np.random.Generator(np.random.PCG64(seed))
    rank = {}
    with open(path, "w") as f:
        for q, l in zip(qids, lists):
            scored = [(int(p), float(s)) for i, (p, s) in enumerate(zip(l, rng.standard_normal(len(l)))) if i % drop_every]
            scored.sort(key=lambda x: x[1], reverse=True)
            rank[int(q)] = [p for p, _ in scored]
            for i, (p, s) in enumerate(scored):
                f.write(f"{int(q)}\t{p}\t{i + 1}\t{s}\n")
    return rank


def test_teacher_rerank_orders_the_candidates_before_the_cut(tmp_path):
    """f-4: groups are cut from the TEACHER's order of the student's top-200 (the reference's re-ranker output),
    not from the student's own ranks."""
    qids, lists = _ranked(nq=6)
    teacher = _teacher_file(tmp_path / "teacher.run", qids[:5], lists[:5])       # the last query has no teacher scores
    t_qids, t_lists = CU.read_run(tmp_path / "teacher.run")
    rer = CU.rerank_with_teacher(qids, lists, t_qids, t_lists)
    for q, l, r in zip(qids[:5], lists[:5], rer[:5]):
        t = teacher[int(q)]
        assert r[:len(t)].tolist() == t                                          # teacher order first
        unscored = [int(p) for i, p in enumerate(l) if i % 7 == 0]
        assert r[len(t):].tolist() == unscored                                   # then the rest in student order
        assert sorted(r.tolist()) == sorted(l.tolist())
    assert rer[5].tolist() == lists[5].tolist()
    dropped = CU.rerank_with_teacher(qids, lists, t_qids, t_lists, keep_unscored=False)
    assert dropped[0].tolist() == teacher[int(qids[0])]
    ex = CU.build_groups(qids, rer, seed=0)
    assert ex[0]["relT_pids"] == teacher[int(qids[0])][:10]
    # through the CLI
    run = tmp_path / "student.run"
    with open(run, "w") as f:
        for q, l in zip(qids, lists):
            for i, p in enumerate(l):
                f.write(f"{int(q)}\t{int(p)}\t{i + 1}\t{200.0 - i}\n")
    spec = importlib.util.spec_from_file_location("mk2", os.path.join(ROOT, "cl-drd_b200", "retriever", "make_curriculum_groups.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    out = tmp_path / "groups.json"
    mk.main(mk.get_args(["--run_path", str(run), "--teacher_run_path", str(tmp_path / "teacher.run"), "--output_path", str(out),
                         "--label_mode", "9"]))
    got = [json.loads(ln) for ln in out.read_text().splitlines()]
    assert got == ex


def _py_read_run(path):
    """The reference's reader loop (evaluation/retrieval_evaluator.py:46-63), lists per qid in first-seen order."""
    order, lists = {}, []
    with open(path, "r") as f:
        for line in f:
            a = line.strip().split("\t")
            if not 2 <= len(a) <= 4:
                raise ValueError("array length is not legal.")
            q, p = int(a[0]), int(a[1])
            if q not in order:
                order[q] = len(lists)
                lists.append([])
            lists[order[q]].append(p)
    return list(order), lists


def _same(native, ref):
    q, lists = native
    return q.tolist() == ref[0] and len(lists) == len(ref[1]) and all(a.tolist() == b for a, b in zip(lists, ref[1]))


def test_native_run_reader_equals_the_reference_reader_loop(tmp_path):
    """cldrd_read_run (parallel, two passes) against the line loop of the reference's reader: column variants, CRLF,
    blanks around fields, an unterminated last line, qids that come back later, int64 extremes, the empty file."""
    p = tmp_path / "a.run"
    p.write_bytes(b"7\t100\n7\t101\t2\n  8\t200\t1\t3.5\r\n9\t-1\t1\t-3.4028234663852886e+38\n7\t102\t3\t1.0  \n"
                  b"+10\t 300 \t1\n-9223372036854775808\t9223372036854775807\t1\t0.5\n8\t201")
    assert _same(CU.read_run(p), _py_read_run(p))
    q, lists = CU.read_run(p)
    assert q.tolist() == [7, 8, 9, 10, -2**63] and lists[0].tolist() == [100, 101, 102] and lists[1].tolist() == [200, 201]
    (tmp_path / "empty.run").write_bytes(b"")
    q, lists = CU.read_run(tmp_path / "empty.run")
    assert q.shape == (0,) and lists == []
    for bad, what in ((b"1\t2\n3\n", "array length"), (b"1\t2\t3\t4\t5\n", "array length"), (b"1\t2\n\n3\t4\n", "array length"),
                      (b"1\tx\n", "integer"), (b"1.5\t2\n", "integer"), (b"1\t2\n \n", "array length")):
        (tmp_path / "bad.run").write_bytes(bad)
        with pytest.raises(ValueError, match=what):
            CU.read_run(tmp_path / "bad.run")
        with pytest.raises(ValueError):
            _py_read_run(tmp_path / "bad.run")
    (tmp_path / "bad.run").write_bytes(b"1\t2\n3\t99999999999999999999\n")       # an id no int64 array can hold
    with pytest.raises(ValueError, match="integer"):
        CU.read_run(tmp_path / "bad.run")
    with pytest.raises(Exception):
        CU.read_run(tmp_path / "missing.run")


@pytest.mark.parametrize("threads", [1, 3, 8])
def test_native_run_reader_across_thread_ranges(tmp_path, threads):
    """A file of several 4 MiB ranges (the unit a reader thread gets), lines of uneven length so that range borders fall
    inside lines, with and without a final newline: same arrays for every thread count, equal to the reference loop."""
    import cldrd
    rng = np.random.default_rng(threads)
    nq, k = 1500, 200
    D = (rng.standard_normal((nq, k)) * 10.0 ** rng.integers(-3, 6, (nq, 1))).astype(np.float32)
    I = rng.integers(0, 2**40, (nq, k), dtype=np.int64) // rng.integers(1, 2**30, (nq, k))
    qids = rng.permutation(nq).astype(np.int64) * 977 + 3
    run = tmp_path / "big.run"
    cldrd.write_run_file(str(run), qids, I, D)
    assert os.path.getsize(run) > 9 << 20
    q, p = CU.read_run_arrays(run, threads)
    assert np.array_equal(q, np.repeat(qids, k)) and np.array_equal(p, I.reshape(-1))
    raw = run.read_bytes()
    run.write_bytes(raw[:-1])                               # no newline at the end of the file
    q2, p2 = CU.read_run_arrays(run, threads)
    assert np.array_equal(q2, q) and np.array_equal(p2, p)
    if threads == 3:
        assert _same(CU.read_run(run, threads), _py_read_run(run))


def test_matrix_path_and_one_by_one_path_cut_the_same_groups():
    """build_groups cuts plain queries (no padding, no repeats, no positive, common length) as one matrix per block and
    the rest one by one; both follow the same rule with the same keys.  Ranks beyond every window, appended in varying
    numbers, push queries onto the one-by-one path without changing what they should get."""
    qids, lists = _ranked(nq=300, depth=200, seed=9)
    fast = CU.build_groups(qids, lists, seed=4)
    longer = [np.concatenate([l, np.arange(10**6, 10**6 + 1 + (i % 7))]) if i % 3 else l for i, l in enumerate(lists)]
    assert fast == CU.build_groups(qids, longer, seed=4)
    # padding behind rank 200 and a repeated pid behind rank 200 change nothing either
    messy = [np.concatenate([l, [-1, -1, l[0]]]) if i % 2 else l for i, l in enumerate(lists)]
    assert fast == CU.build_groups(qids, messy, seed=4)
    # qrels for a few queries: the others keep their draws
    qrels = {int(qids[5]): [int(lists[5][77])], int(qids[6]): []}
    with_q = CU.build_groups(qids, lists, seed=4, qrels=qrels)
    assert [e for i, e in enumerate(with_q) if i != 5] == [e for i, e in enumerate(fast) if i != 5]
    assert with_q[5]["relT_pids"][0] == int(lists[5][77])
    # all sizes at their limits: windows exactly as wide as the draw
    e = CU.build_groups(qids[:3], lists[:3], n_rel=5, n_most=5, n_semi=190, most_window=(5, 10), semi_window=(10, 200))
    assert all(x["most_hard_pids"] == l[5:10].tolist() and x["semi_hard_pids"] == l[10:200].tolist() for x, l in zip(e, lists))


def test_write_groups_text_is_json_dumps_text(tmp_path):
    qids, lists = _ranked(nq=40)
    ex = CU.groups_for_label_mode(qids, lists, "7", seed=3)
    ex.append({"qid": -5, "relT_pids": [], "most_hard_pids": [2**62], "semi_hard_pids": [-1, 0]})
    ex.append({"qid": 7, "relT_pids": [1], "most_hard_pids": [2], "semi_hard_pids": [3], "extra": "x"})      # not the plain shape
    ex.append({"qid": np.int64(8), "relT_pids": [1], "most_hard_pids": [2], "semi_hard_pids": [3]})           # numpy scalar
    out = tmp_path / "g.json"
    ok = [e for e in ex if not isinstance(e["qid"], np.integer)]
    assert CU.write_groups(out, ok) == len(ok)
    assert out.read_text() == "".join(json.dumps(e) + "\n" for e in ok)
    with pytest.raises(TypeError):
        CU.write_groups(out, ex[-1:])          # json.dumps refuses numpy scalars: no silent change of behaviour
