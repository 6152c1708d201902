"""CPU tests of the oracle against the committed golden vectors (oracle/flat_ip.py).
PARITY UNPINNED upstream: see the oracle header."""
import os
import struct

import numpy as np

from oracle import flat_ip as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")

# hand-assembled IxMp{IxFI}: 2 vectors, d=4, ids (7, 2**33+5)  -- SURVEY §8 a-2 byte offsets
TINY = (
    b"IxMp" + struct.pack("<i", 4) + struct.pack("<q", 2) + struct.pack("<q", 1 << 20) * 2 + b"\x01" + struct.pack("<i", 0)
    + b"IxFI" + struct.pack("<i", 4) + struct.pack("<q", 2) + struct.pack("<q", 1 << 20) * 2 + b"\x01" + struct.pack("<i", 0)
    + struct.pack("<Q", 8) + struct.pack("<8f", 1.0, 2.0, 3.0, 4.0, -1.0, 0.5, 0.25, 8.0)
    + struct.pack("<Q", 2) + struct.pack("<2q", 7, 2 ** 33 + 5)
)


def test_tiny_file_layout_offsets():
    assert len(TINY) == 82 + 4 * 2 * 4 + 8 + 8 * 2
    assert TINY[0:4] == b"IxMp" and TINY[37:41] == b"IxFI"
    assert struct.unpack_from("<Q", TINY, 74)[0] == 8          # count = ntotal*d at byte 74
    assert struct.unpack_from("<f", TINY, 82)[0] == 1.0        # payload starts at byte 82
    with open(os.path.join(GOLD, "tiny_ixmp.index"), "rb") as f:
        assert f.read() == TINY


def test_tiny_file_reader_writer():
    xb, ids, info = O.read_index_bytes(TINY)
    assert info["d"] == 4 and info["ntotal"] == 2 and info["data_off"] == 82 and info["metric"] == 0
    assert xb.tolist() == [[1.0, 2.0, 3.0, 4.0], [-1.0, 0.5, 0.25, 8.0]]
    assert ids.tolist() == [7, 2 ** 33 + 5]
    assert O.write_index_bytes(xb, ids) == TINY
    assert O.write_index_bytes(xb, ids, idmap2=True)[:4] == b"IxM2"
    bare = O.write_index_bytes(xb, None)
    assert bare[:4] == b"IxFI" and len(bare) == 45 + 32
    xb2, ids2, _ = O.read_index_bytes(bare)
    assert ids2 is None and np.array_equal(xb2, xb)


def test_msmarco_file_size_formula():
    N, d = 8841823, 768
    assert 82 + 4 * N * d + 8 + 8 * N == 27232814930


def test_search_tiny_known_answer():
    xb, ids, _ = O.read_index_bytes(TINY)
    q = np.array([[1, 0, 0, 0], [0, 0, 0, 1]], dtype=np.float32)
    D, I = O.search(xb, ids, q, 3)
    assert D[0].tolist()[:2] == [1.0, -1.0] and I[0].tolist() == [7, 2 ** 33 + 5, -1]
    assert D[1].tolist()[:2] == [8.0, 4.0] and I[1].tolist() == [2 ** 33 + 5, 7, -1]
    assert D[0, 2] == O.NEG_FLT_MAX


def test_seeded_golden_1000x64():
    g = np.load(os.path.join(GOLD, "seeded_1000x64.npz"))
    xb, xq, ids = O.synth(1000, 64, 0), O.synth(16, 64, 1), O.synth_ids(1000, 7)
    for k in (10, 100):
        D, I = O.search(xb, ids, xq, k)
        assert np.array_equal(I, g[f"I{k}"])
        np.testing.assert_allclose(D, g[f"D{k}"], rtol=1e-6)
        # block-merged walk gives the same answer as the single-block walk
        D2, R2 = O.search_rows(xb, xq, k, block=137)
        assert np.array_equal(ids[R2], I)


def test_ties_lower_row_first_and_padding():
    xb = np.ones((5, 4), dtype=np.float32)
    xb[3] = 2.0
    q = np.ones((1, 4), dtype=np.float32)
    D, R = O.search_rows(xb, q, 8)
    assert R[0].tolist() == [3, 0, 1, 2, 4, -1, -1, -1]
    assert D[0, 5] == O.NEG_FLT_MAX
    D, R = O.search_rows(xb, q, 3, block=2)
    assert R[0].tolist() == [3, 0, 1]


def test_fp64_twin_agrees_with_fp32_within_tolerance():
    xb, xq = O.synth(3000, 768, 3), O.synth(4, 768, 4)
    D32, R32 = O.search_rows(xb, xq, 50)
    D64, R64 = O.search_rows(xb, xq, 51, dtype=np.float64)
    r = O.compare_topk(D32, R32, D64[:, :50].astype(np.float32), R64[:, :50], D64, R64)
    assert r["ok"], r


def test_compare_topk_flags_real_differences():
    xb, xq = O.synth(500, 32, 5), O.synth(3, 32, 6)
    D, R = O.search_rows(xb, xq, 20)
    assert O.compare_topk(D, R, D, R)["ok"]
    R2 = R.copy()
    R2[0, 0], R2[0, 5] = R2[0, 5], R2[0, 0]
    assert not O.compare_topk(D, R2, D, R)["ok"]
    D2 = D.copy()
    D2[1, 3] *= 1.001
    assert not O.compare_topk(D2, R, D, R)["ok"]


def test_index_retrieve_shapes():
    xb, xq, ids = O.synth(300, 16, 0), O.synth(10, 16, 1), O.synth_ids(300)
    D, I = O.index_retrieve(xb, ids, xq, 5)
    assert isinstance(D, np.ndarray) and D.shape == (10, 5)
    s, nn = O.index_retrieve(xb, ids, xq, 5, batch=4)
    assert isinstance(s, list) and len(s) == 10 and len(nn[0]) == 5
    assert nn == I.tolist()


def test_run_writer_golden(tmp_path):
    Dr = np.array([[103.856, 71.5, 0.1, -2.25e-5], [1e16, 3.0, 1.5e-7, -0.0]], dtype=np.float32)
    Ir = np.array([[5, 2 ** 33 + 5, 0, -1], [9, 8, 7, 6]], dtype=np.int64)
    p = tmp_path / "run.tsv"
    avg = O.write_run(str(p), [1048585, 2], Ir, Dr)
    assert avg == 4.0
    with open(os.path.join(GOLD, "run_golden.tsv")) as f:
        assert p.read_text() == f.read()
    assert p.read_text().splitlines()[0] == "1048585\t5\t1\t103.85600280761719"


def test_faiss_made_fixtures():
    """Consumes fixtures written by tests/golden/make_golden_faiss.py on a machine with real faiss (byte-equal index
    file, search results inside the parity rule, the -1 / -FLT_MAX padding).  faiss is absent from this image and
    un-pinned upstream, so until somebody commits those files this test reports PARITY UNPINNED and skips."""
    import pytest
    need = ["faiss_1000x64.index", "faiss_1000x64.npz", "faiss_20000x768_k1000.npz", "faiss_padding.npz"]
    if not all(os.path.exists(os.path.join(GOLD, f)) for f in need):
        pytest.skip("PARITY UNPINNED: no faiss-made fixtures under tests/golden/ (run tests/golden/make_golden_faiss.py where faiss exists)")
    xb, xq, ids = O.synth(1000, 64, 0), O.synth(16, 64, 1), O.synth_ids(1000, 7)
    with open(os.path.join(GOLD, "faiss_1000x64.index"), "rb") as f:
        assert f.read() == O.write_index_bytes(xb, ids)
    g = np.load(os.path.join(GOLD, "faiss_1000x64.npz"))
    for k in (10, 100):
        D, I = O.search(xb, ids, xq, k)
        r = O.compare_topk(D, I, g[f"D{k}"], g[f"I{k}"], *O.search(xb, ids, xq, k + 16, dtype=np.float64))
        assert r["ok"], (k, r)
    g = np.load(os.path.join(GOLD, "faiss_20000x768_k1000.npz"))
    xb, xq = O.synth(20000, 768, 0), O.synth(8, 768, 1)
    D, R = O.search_rows(xb, xq, 1000)
    r = O.compare_topk(D, R.astype(np.int64), g["D"], g["R"].astype(np.int64), *O.search_rows(xb, xq, 1016, dtype=np.float64))
    assert r["ok"], r
    g = np.load(os.path.join(GOLD, "faiss_padding.npz"))
    Dp, Ip = O.search(O.synth(5, 8, 2), None, O.synth(3, 8, 3), 8)
    assert np.array_equal(Ip[:, 5:], g["I"][:, 5:]) and np.array_equal(Dp[:, 5:], g["D"][:, 5:])
    assert np.array_equal(Ip, g["I"])


def test_second_cpu_implementation_agrees_with_the_oracle():
    """Comparator matrix: the numpy oracle, its fp64 twin and an independently written CPU search (torch-CPU sgemm +
    topk, blocked over rows and merged -- oracle/cpu_baseline.py) agree inside the parity rule on seeded inputs, so a
    bug would have to be made twice, in different code, to go unnoticed."""
    import torch
    from oracle import cpu_baseline as CB
    xb, xq = O.synth(30000, 96, 40), O.synth(37, 96, 41)
    k = 200
    D, R = O.search_rows(xb, xq, k)
    D2, R2 = CB.search_torch_cpu(torch.from_numpy(xb), torch.from_numpy(xq), k, batch=16, row_block=4096)
    ext = O.search_rows(xb, xq, k + 16, dtype=np.float64)
    r = O.compare_topk(D2.numpy(), R2.numpy().astype(np.int64), D, R.astype(np.int64), *ext)
    assert r["ok"] and r["overlap"] == 1.0, r
    D64, R64 = O.search_rows(xb, xq, k, dtype=np.float64)
    r = O.compare_topk(D64.astype(np.float32), R64.astype(np.int64), D, R.astype(np.int64), *ext)
    assert r["ok"], r


def test_bench_checker_agrees_with_the_oracle():
    """The brute-force checker behind bench.py's `parity` block and the full-size GPU test (fp64 torch matmul over a
    shard's rows, stable sort: score descending, row ascending) is pinned here, on the CPU, to the oracle's fp64 search:
    same rows in the same order, including exact duplicates (ties -> lower row) and fewer rows than asked for."""
    import torch
    import bench
    xb, xq = O.synth(7000, 48, 50), O.synth(9, 48, 51)
    xb[100:110] = xb[5]                                   # exact duplicates: ties
    s, r = bench.brute_force_shard(torch, torch.from_numpy(xb), 1000, torch.from_numpy(xq), 64)
    D64, R64 = O.search_rows(xb, xq, 64, dtype=np.float64)
    assert np.array_equal(r.numpy() - 1000, R64)
    assert np.allclose(s.numpy(), D64, rtol=1e-12, atol=0)
    s, r = bench.brute_force_shard(torch, torch.from_numpy(xb[:5]), 0, torch.from_numpy(xq), 8)
    assert (r.numpy()[:, 5:] == -1).all() and np.isinf(s.numpy()[:, 5:]).all()


def test_oracle_restatements_reproduce_the_reference_run_here(tmp_path):
    """The oracle's restatements of the reference's OWN code (index_retrieve loop, regroup, f-string writer, meta.pkl)
    against files the reference's scripts produced in this container (tests/golden/ref_pipeline/, generator:
    tests/golden/make_golden_reference.py): byte for byte.  This pins those restatements to the reference itself; the
    search arithmetic and the index-file bytes inside the fixture are the oracle's (faiss is absent) and stay unpinned."""
    import gzip
    import pickle
    fix = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_pipeline")
    xb, ids, _ = O.read_index(os.path.join(fix, "checkpoint_120000.index"))
    for run, embs, tsv, k in (("dev.run.gz", "query_embs.npy", "queries.dev.tsv", 1000),
                              ("passages.run.gz", "passage_embs.npy", "passages.small.tsv", 200)):
        xq = np.load(os.path.join(fix, embs))
        text_ids = [int(ln.split("\t")[0]) for ln in open(os.path.join(fix, tsv))]
        nn_scores, nn_ids = O.index_retrieve(xb, ids, xq, k, batch=128)
        assert isinstance(nn_scores, list) and isinstance(nn_scores[0][0], float)
        out = tmp_path / run[:-3]
        avg = O.write_run(str(out), text_ids, nn_ids, nn_scores)
        assert out.read_bytes() == gzip.open(os.path.join(fix, run), "rb").read()
        assert avg == float(k)
    O.write_meta(str(tmp_path), ids.tolist())
    ours, gold = (pickle.load(open(os.path.join(d, "meta.pkl"), "rb")) for d in (str(tmp_path), fix))
    assert ours["text_ids"].dtype == gold["text_ids"].dtype and ours["text_ids"].tolist() == gold["text_ids"].tolist()
    assert ours["text_id_to_idx"] == gold["text_id_to_idx"] and list(ours["text_id_to_idx"]) == list(gold["text_id_to_idx"])
    assert open(os.path.join(str(tmp_path), "meta.pkl"), "rb").read() == open(os.path.join(fix, "meta.pkl"), "rb").read()


def test_compare_topk_cannot_be_fooled_by_plausible_lists():
    """The parity checker is what every GPU result is judged by: lists that look right but are not must fail --
    a repeated id, a foreign id carrying the right score, a swap across the k boundary outside the near-tie band,
    wrong padding -- and the allowances must stay as narrow as the rule says (ties as sets, boundary band only)."""
    xb, xq = O.synth(400, 16, 8), O.synth(4, 16, 9)
    xb[3] = xq[0] * 0.5                                      # near the top for query 0 ...
    xb[7] = xb[3]                                            # ... and rows 3 and 7 tie exactly for every query
    k = 30
    D, R = O.search_rows(xb, xq, k)
    De, Re = O.search_rows(xb, xq, k + 8, dtype=np.float64)
    assert O.compare_topk(D, R, D, R, De, Re)["ok"]
    # a tie pair may come in either order ...
    assert {3, 7} <= set(R[0].tolist())
    for i in range(4):
        pos = {int(r): j for j, r in enumerate(R[i])}
        if 3 in pos and 7 in pos:
            R2 = R.copy()
            R2[i, pos[3]], R2[i, pos[7]] = 7, 3
            assert O.compare_topk(D, R2, D, R, De, Re)["ok"]
            # ... but not twice the same member
            R3 = R.copy()
            R3[i, pos[7]] = 3
            assert not O.compare_topk(D, R3, D, R, De, Re)["ok"]
    # a foreign row with the right score in its place
    R4 = R.copy()
    R4[0, 4] = int(Re[0, k + 5])
    r = O.compare_topk(D, R4, D, R, De, Re)
    assert not r["ok"] and r["bad_ids"] >= 1 and r["overlap"] < 1.0
    # the k-th row exchanged for the (k+1)-th: excused only when fp64 says they are within the tolerance (they are not here)
    R5, D5 = R.copy(), D.copy()
    R5[1, k - 1] = int(Re[1, k])
    D5[1, k - 1] = np.float32(De[1, k])
    assert abs(De[1, k] - De[1, k - 1]) > 1e-5 * abs(De[1, k - 1])
    assert not O.compare_topk(D5, R5, D, R, De, Re)["ok"]
    # padding must match to the bit
    Dp, Rp = O.search_rows(xb[:20], xq, k)
    assert (Rp[:, 20:] == -1).all() and O.compare_topk(Dp, Rp, Dp, Rp)["ok"]
    Rq = Rp.copy()
    Rq[0, 25] = 0
    assert not O.compare_topk(Dp, Rq, Dp, Rp)["ok"]
    Dq = Dp.copy()
    Dq[0, 25] = 0.0
    assert not O.compare_topk(Dq, Rp, Dp, Rp)["ok"]
    # scores: 1e-5 relative is the limit, on every position
    D6 = D.copy()
    D6[2, 11] = D6[2, 11] * np.float32(1 + 3e-5)
    assert not O.compare_topk(D6, R, D, R, De, Re)["ok"]
    D7 = D.copy()
    D7[2, 11] = np.nextafter(D7[2, 11], np.float32(np.inf))
    assert O.compare_topk(D7, R, D, R, De, Re)["ok"]
