"""Host-side logic of the faiss-shaped surface: index objects, file I/O through the C ABI,
the run-file writer, and the reference's own evaluator consuming our output."""
import os
import sys

import numpy as np
import pytest

from oracle import flat_ip as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")
REF = "/root/reference"


def test_write_index_matches_oracle_bytes(cldrd_lib, tmp_path):
    import cldrd
    xb, ids = O.synth(257, 24, 0), O.synth_ids(257)
    idx = cldrd.IndexIDMap(cldrd.IndexFlatIP(24))
    idx.add_with_ids(xb[:100], ids[:100])
    idx.add_with_ids(xb[100:], ids[100:])
    assert idx.ntotal == 257 and idx.d == 24
    p = tmp_path / "a.index"
    cldrd.write_index(idx, str(p))
    assert p.read_bytes() == O.write_index_bytes(xb, ids)
    idx2 = cldrd.IndexIDMap2(cldrd.IndexFlatIP(24))
    idx2.add_with_ids(xb, ids)
    cldrd.write_index(idx2, str(p))
    assert p.read_bytes() == O.write_index_bytes(xb, ids, idmap2=True)
    flat = cldrd.index_factory(24, "Flat", cldrd.METRIC_INNER_PRODUCT)
    flat.add(xb)
    cldrd.write_index(flat, str(p))
    assert p.read_bytes() == O.write_index_bytes(xb, None)


def test_read_index_golden_and_lazy(cldrd_lib):
    import cldrd
    idx = cldrd.read_index(os.path.join(GOLD, "tiny_ixmp.index"))
    assert isinstance(idx, cldrd.IndexIDMap) and idx.ntotal == 2 and idx.d == 4
    assert idx._rows.file is not None  # rows not read yet
    assert idx.id_map.tolist() == [7, 2 ** 33 + 5]
    assert idx._rows.materialize().tolist() == [[1.0, 2.0, 3.0, 4.0], [-1.0, 0.5, 0.25, 8.0]]


def test_streaming_writer_and_row_reads(cldrd_lib, tmp_path):
    import ctypes as C
    from cldrd._lib import check, ptr
    xb, ids = O.synth(1000, 16, 2), O.synth_ids(1000)
    p = str(tmp_path / "s.index").encode()
    w = C.c_void_p()
    check(cldrd_lib.cldrd_index_writer_begin(C.byref(w), p, 1000, 16, 1, 0))
    for r0 in range(0, 1000, 300):
        part = np.ascontiguousarray(xb[r0:r0 + 300])
        check(cldrd_lib.cldrd_index_writer_append(w, ptr(part), part.shape[0]))
    check(cldrd_lib.cldrd_index_writer_finish(w, ptr(ids)))
    assert open(p, "rb").read() == O.write_index_bytes(xb, ids)
    out = np.empty((10, 16), dtype=np.float32)
    check(cldrd_lib.cldrd_index_read_rows(p, 495, 10, ptr(out)))
    assert np.array_equal(out, xb[495:505])
    oi = np.empty((7,), dtype=np.int64)
    check(cldrd_lib.cldrd_index_read_ids(p, 993, 7, ptr(oi)))
    assert np.array_equal(oi, ids[993:])
    assert cldrd_lib.cldrd_index_read_rows(p, 995, 10, ptr(out)) != 0


def test_bad_files_rejected(cldrd_lib, tmp_path):
    import cldrd
    p = tmp_path / "bad.index"
    p.write_bytes(b"IwFl" + b"\0" * 100)
    with pytest.raises(cldrd.CldrdError):
        cldrd.read_index(str(p))
    good = O.write_index_bytes(O.synth(4, 4, 0), O.synth_ids(4))
    p.write_bytes(good[:-5])
    with pytest.raises(cldrd.CldrdError):
        cldrd.read_index(str(p))
    l2 = bytearray(good)
    l2[70:74] = (1).to_bytes(4, "little")  # inner metric = L2
    p.write_bytes(bytes(l2))
    with pytest.raises(cldrd.CldrdError):
        cldrd.read_index(str(p))


def test_dtype_and_shape_errors_like_faiss(cldrd_lib):
    import cldrd
    idx = cldrd.IndexIDMap(cldrd.IndexFlatIP(8))
    with pytest.raises(TypeError):
        idx.add_with_ids(np.zeros((2, 8), dtype=np.float64), np.arange(2, dtype=np.int64))
    with pytest.raises(TypeError):
        idx.add_with_ids(np.zeros((2, 8), dtype=np.float32), np.arange(2, dtype=np.int32))
    with pytest.raises(AssertionError):
        idx.add_with_ids(np.zeros((2, 7), dtype=np.float32), np.arange(2, dtype=np.int64))
    with pytest.raises(AssertionError):
        idx.add_with_ids(np.zeros((2, 8), dtype=np.float32), np.arange(3, dtype=np.int64))


def test_run_writer_bytes_equal_reference_loop(cldrd_lib, tmp_path):
    import cldrd
    rng = np.random.default_rng(0)
    n, k = 37, 50
    D = (rng.standard_normal((n, k)) * np.array([1e-6, 1e-3, 1, 100, 1e7])[rng.integers(0, 5, (n, k))]).astype(np.float32)
    D[0, :5] = [0.0, -0.0, 1e16, 9.999e15, 1e-4]
    I = rng.integers(-1, 2 ** 40, (n, k), dtype=np.int64)
    qids = rng.permutation(10 ** 6)[:n].astype(np.int64)
    a, b = tmp_path / "a.tsv", tmp_path / "b.tsv"
    avg_a = cldrd.write_run_file(str(a), qids, I, D)
    avg_b = O.write_run(str(b), qids.tolist(), I, D)
    assert a.read_bytes() == b.read_bytes()
    assert avg_a == avg_b == k
    # duplicate qids regroup exactly like the reference's dict
    qd = qids.copy()
    qd[5] = qd[1]
    qd[9] = qd[1]
    avg_a = cldrd.write_run_file(str(a), qd, I, D)
    avg_b = O.write_run(str(b), qd.tolist(), I, D)
    assert a.read_bytes() == b.read_bytes() and avg_a == avg_b
    with open(os.path.join(GOLD, "run_golden.tsv"), "rb") as f:
        gold = f.read()
    Dr = np.array([[103.856, 71.5, 0.1, -2.25e-5], [1e16, 3.0, 1.5e-7, -0.0]], dtype=np.float32)
    Ir = np.array([[5, 2 ** 33 + 5, 0, -1], [9, 8, 7, 6]], dtype=np.int64)
    cldrd.write_run_file(str(a), [1048585, 2], Ir, Dr)
    assert a.read_bytes() == gold


def test_run_writer_threads_do_not_change_the_bytes(cldrd_lib, tmp_path):
    """Many ~2 MiB pieces formatted by 1, 3 and 8 threads, written with pwrite at prefix-summed offsets:
    identical files; equal qids that straddle a piece boundary keep one rank sequence; append continues
    behind the existing bytes (retrieve_top_passages.py:90-109 is the loop all of this replaces)."""
    import cldrd
    rng = np.random.default_rng(5)
    n, k = 3000, 200                     # ~27 MB of text: a dozen pieces
    D = (rng.standard_normal((n, k)) * 30 + 80).astype(np.float32)
    I = rng.integers(0, 8_841_823, (n, k), dtype=np.int64)
    qids = np.repeat(rng.permutation(10 ** 7)[: n // 3].astype(np.int64), 3)[:n]    # runs of three equal qids
    files = []
    for t in (1, 3, 8):
        f = tmp_path / f"t{t}.tsv"
        avg = cldrd.write_run_file(str(f), qids, I, D, threads=t)
        assert avg == 3 * k
        files.append(f.read_bytes())
    assert files[0] == files[1] == files[2]
    ref = tmp_path / "ref.tsv"
    O.write_run(str(ref), qids[:300].tolist(), I[:300], D[:300])
    assert files[0].startswith(ref.read_bytes())
    assert files[0].count(b"\n") == n * k
    # append in two halves == one call (the halves are cut between two qid runs)
    f2 = tmp_path / "halves.tsv"
    cldrd.write_run_file(str(f2), qids[:1500], I[:1500], D[:1500], threads=4)
    cldrd.write_run_file(str(f2), qids[1500:], I[1500:], D[1500:], append=True, threads=4)
    assert f2.read_bytes() == files[0]


def test_run_file_stream_equals_one_call(cldrd_lib, tmp_path):
    """RunFileStream (blocks appended by a writer thread while the caller goes on searching) == one write_run_file
    call over the concatenated blocks == the reference's loop (retrieve_top_passages.py:90-109)."""
    import cldrd
    rng = np.random.default_rng(9)
    n, k = 5000, 37
    D = (rng.standard_normal((n, k)) * 20 + 60).astype(np.float32)
    I = rng.integers(0, 8_841_823, (n, k), dtype=np.int64)
    qids = rng.permutation(10 ** 7)[:n].astype(np.int64)
    a, b = tmp_path / "sub" / "stream.tsv", tmp_path / "one.tsv"
    st = cldrd.RunFileStream(str(a))             # creates the parent directory like the reference (:99-100)
    for c0 in range(0, n, 777):
        st.put(qids[c0:c0 + 777], I[c0:c0 + 777], D[c0:c0 + 777])
    assert st.close() == k
    assert cldrd.write_run_file(str(b), qids, I, D) == k
    assert a.read_bytes() == b.read_bytes()
    ref = tmp_path / "ref.tsv"
    O.write_run(str(ref), qids[:200].tolist(), I[:200], D[:200])
    assert a.read_bytes().startswith(ref.read_bytes())


def test_score_text_matches_python_repr(cldrd_lib):
    import cldrd
    rng = np.random.default_rng(1)
    vals = np.concatenate([
        rng.standard_normal(2000).astype(np.float32) * 100,
        (10.0 ** rng.uniform(-12, 20, 2000)).astype(np.float32),
        np.array([0, -0.0, 1, 16777216, 1e16, 9.9999998e15, 1e-4, 9.9999e-5, 3.4028235e38, 1.4e-45], dtype=np.float32),
    ])
    for v in vals:
        assert cldrd.format_score(v) == repr(float(v)), float(v)


def test_score_text_fast_path_equals_general_routine(cldrd_lib):
    """The writer prints fp32 scores through a specialised exact routine (csrc/runfile.cpp: shortest_f32); the library
    compares it with the general std::to_chars routine itself.  Here: 2^22 consecutive patterns around typical scores,
    a stride that visits every exponent of both signs, and the exact-tie / power-of-two cases against Python's repr.
    All 2^32 patterns: tools/check_score_text.py (profiles/r02_score_text_sweep.json: 0 differences)."""
    import ctypes as C
    import cldrd
    bad, fast = C.c_uint32(0), C.c_int64(0)
    first = int(np.array([100.0], dtype=np.float32).view(np.uint32)[0])
    assert cldrd_lib.cldrd_format_score_selfcheck(first, 1, 1 << 22, C.byref(bad), C.byref(fast)) == 0, hex(bad.value)
    assert fast.value == 1 << 22                       # these scores never leave the fast path
    assert cldrd_lib.cldrd_format_score_selfcheck(12345, 1021, 1 << 22, C.byref(bad), C.byref(fast)) == 0, hex(bad.value)
    assert 0 < fast.value < 1 << 22                    # zero / denormals / huge / inf / nan take the general one
    ties = [512 + 1 / 2**14, 512 + 3 / 2**14, 1024 + 5 / 2**14, 2.0**-33, 2.0**-34, 2.0**52, 2.0**53, 2.0**24 - 1,
            0.1, 1e-4, 9.9999e-5, 1e16, 9.9999998e15, 8388608.5, 0.0078125, 3.0e-10, 1.17549435e-38]
    for v in ties + [-t for t in ties]:
        v32 = np.float32(v)
        assert cldrd.format_score(v32) == repr(float(v32)), float(v32)
    for e in range(1, 255):                            # every power of two and its two neighbours
        for frac in (0, 1, 0x7fffff):
            v32 = np.array([(e << 23) | frac], dtype=np.uint32).view(np.float32)[0]
            assert cldrd.format_score(v32) == repr(float(v32)), float(v32)


def test_run_writer_extreme_ids_and_rank_carries(cldrd_lib, tmp_path):
    """int64 extremes in every integer column, -1 pads, rank sequences that cross 9 -> 10 -> 100 -> 1000 and a
    duplicate qid whose sequence continues past 99 999 -> 100 000: same bytes as the reference's loop."""
    import cldrd
    import oracle.flat_ip as O
    k = 1200
    rng = np.random.default_rng(5)
    D = -np.sort(-rng.standard_normal((4, k)).astype(np.float32) * 1e3, axis=1)
    I = rng.integers(-2**63, 2**63 - 1, (4, k), dtype=np.int64)
    I[0, :3] = [-1, 2**63 - 1, -2**63]
    D[1, -2:] = np.float32(-3.4028235e38)
    qids = np.array([-2**63, 2**63 - 1, 7, 7], dtype=np.int64)
    a, b = tmp_path / "a.tsv", tmp_path / "b.tsv"
    cldrd.write_run_file(str(a), qids, I, D, threads=2)
    O.write_run(str(b), qids.tolist(), I, D)
    assert a.read_bytes() == b.read_bytes()
    # 84 rows of one qid: its rank column runs from 1 to 100 800
    qd = np.full(84, 3, dtype=np.int64)
    D2 = np.tile(D[:1], (84, 1))
    I2 = np.tile(I[1:2], (84, 1))
    cldrd.write_run_file(str(a), qd, I2, D2, threads=3)
    O.write_run(str(b), qd.tolist(), I2, D2)
    assert a.read_bytes() == b.read_bytes()
    assert a.read_bytes().splitlines()[-1].split(b"\t")[2] == b"100800"


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present on this box")
def test_reference_evaluator_consumes_our_run_file(cldrd_lib, tmp_path):
    """evaluation/retrieval_evaluator.py:42-76 reads cols 0,1 in file order: feed it our writer's output."""
    import importlib.util
    import cldrd
    spec = importlib.util.spec_from_file_location("ref_eval", os.path.join(REF, "evaluation", "retrieval_evaluator.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    xb, xq, ids = O.synth(2000, 32, 0), O.synth(20, 32, 1), O.synth_ids(2000)
    D, I = O.search(xb, ids, xq, 100)
    qids = np.arange(100, 120, dtype=np.int64)
    run = tmp_path / "runs" / "dev.run"
    cldrd.write_run_file(str(run), qids, I, D)
    qrels = tmp_path / "qrels.tsv"
    with open(qrels, "w") as f:
        for i, q in enumerate(qids):
            f.write(f"{q}\t0\t{I[i, i % 7]}\t1\n")  # the relevant passage sits at rank (i%7)+1
    ev = mod.RankingEvaluator(str(qrels))
    m = ev.compute_metrics(str(run))
    exp = np.mean([1.0 / (i % 7 + 1) for i in range(20)])
    assert abs(m["MRR@10"] - exp) < 1e-9 and m["QueriesRanked"] == 20


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present on this box")
def test_reference_retrieval_utils_imports_against_our_faiss(cldrd_lib):
    """retriever/retrieval_utils.py does `import faiss`; with cl-drd_b200/compat on the path it
    gets ours, and its own convert_index_to_gpu / index_retrieve can drive our objects."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import sys; sys.path[:0]=[%r,%r]\n"
        "import faiss, retriever.retrieval_utils as ru\n"
        "assert ru.__file__.startswith(%r), ru.__file__\n"
        "assert faiss.__version__=='cldrd-b200'\n"
        "assert callable(ru.index_retrieve) and callable(ru.convert_index_to_gpu)\n"
        "idx=faiss.IndexIDMap(faiss.IndexFlatIP(8)); print('ok', idx.ntotal)\n"
    ) % (os.path.join(root, "cl-drd_b200", "compat"), REF, REF)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=REF, timeout=300)
    assert r.returncode == 0 and "ok 0" in r.stdout, r.stderr[-2000:]


def test_run_writer_random_shapes_equal_reference_loop(cldrd_lib, tmp_path):
    """Random (nq, k, thread count, duplicate-qid runs, score scales from 1e-12 to 1e12, append splits): always the bytes
    of the reference's regroup + f-string loops (retrieve_top_passages.py:90-109)."""
    import cldrd
    import oracle.flat_ip as O
    rng = np.random.default_rng(11)
    for case in range(12):
        nq, k = int(rng.integers(1, 400)), int(rng.integers(1, 300))
        scale = 10.0 ** rng.uniform(-12, 12)
        D = (rng.standard_normal((nq, k)) * scale).astype(np.float32)
        if case % 3 == 0:
            D = np.round(D, 2)                         # short decimals: the formatter drops many digits
        I = rng.integers(-1, 2**40, (nq, k), dtype=np.int64)
        qids = np.sort(rng.integers(0, max(2, nq // 2), nq)).astype(np.int64) if case % 2 else rng.permutation(nq).astype(np.int64) + 5
        a, b = tmp_path / f"a{case}.tsv", tmp_path / f"b{case}.tsv"
        cut = int(rng.integers(0, nq + 1))
        while 0 < cut < nq and qids[cut] == qids[cut - 1]:      # an append continues at a qid boundary, like RunFileStream's blocks
            cut += 1
        cldrd.write_run_file(str(a), qids[:cut], I[:cut], D[:cut], threads=int(rng.integers(1, 9)))
        cldrd.write_run_file(str(a), qids[cut:], I[cut:], D[cut:], append=True, threads=int(rng.integers(1, 9)))
        O.write_run(str(b), qids.tolist(), I, D)
        assert a.read_bytes() == b.read_bytes(), (case, nq, k, scale)
