"""Our mirrors against files the REFERENCE'S OWN SCRIPTS produced (tests/golden/ref_pipeline/, made by
tests/golden/make_golden_reference.py running /root/reference/retriever/index_text.py and retrieve_top_passages.py
unmodified on the CPU; only faiss -- absent -- is stood in for by the oracle, see that script's header).

CPU (`-m "not gpu"`): our index build, encoder mirror, retrieve loop and run-file writer reproduce the reference's
index file, meta.pkl, query embeddings and 24 000-line run file (300 padding hits per query included).
GPU: the search on the reference-built index file against the reference-written run under the parity rule.
Nothing here reads /root/reference at run time."""
import contextlib
import gzip
import io
import os
import pickle
import shutil
import sys

import numpy as np
import pytest

from oracle import flat_ip as O

HERE = os.path.dirname(os.path.abspath(__file__))
FIX = os.path.join(HERE, "golden", "ref_pipeline")
sys.path.insert(0, os.path.join(os.path.dirname(HERE), "cl-drd_b200"))

KEEP = ("****", "load ", "# nan", "embs dtype", "retrieve ", "# unique", "average ranks", "Query Num")


def _ref_stdout(section):
    lines = open(os.path.join(FIX, "stdout.txt")).read().splitlines()
    a = lines.index(f"=== {section} ===")
    b = next((i for i in range(a + 1, len(lines)) if lines[i].startswith("===")), len(lines))
    return lines[a + 1:b]


def _golden_run(name="dev.run.gz"):
    return gzip.open(os.path.join(FIX, name), "rb").read()


def _parse_run(text: bytes, k: int):
    rows = [ln.split(b"\t") for ln in text.splitlines()]
    q = np.array([int(r[0]) for r in rows], dtype=np.int64).reshape(-1, k)
    I = np.array([int(r[1]) for r in rows], dtype=np.int64).reshape(-1, k)
    rk = np.array([int(r[2]) for r in rows], dtype=np.int64).reshape(-1, k)
    D = np.array([float(r[3]) for r in rows], dtype=np.float32).reshape(-1, k)
    assert (q == q[:, :1]).all() and (rk == np.arange(1, k + 1)).all()
    return q[:, 0], D, I


def _experiment(tmp_path):
    """The directory shape the reference's guard wants: <exp>/models/<ckpt>, <exp>/index/ (index_text.py:50)."""
    exp = tmp_path / "experiment"
    (exp / "models").mkdir(parents=True)
    shutil.copy(os.path.join(FIX, "checkpoint_120000.pth.tar"), exp / "models")
    return exp


def test_fixture_is_what_the_generator_describes():
    info = __import__("json").load(open(os.path.join(FIX, "README.json")))
    assert info["reference_scripts_run"] == ["retriever/index_text.py", "retriever/retrieve_top_passages.py",
                                             "retriever/retrieve_top_queries.py"]
    for name, size in info["sizes"].items():
        if name != "README.json":
            assert os.path.getsize(os.path.join(FIX, name)) == size, name
    xb, ids, hdr = O.read_index(os.path.join(FIX, "checkpoint_120000.index"))
    meta = pickle.load(open(os.path.join(FIX, "meta.pkl"), "rb"))
    assert hdr["fourcc"] == "IxMp" and xb.shape == (700, 64) and ids.tolist() == meta["text_ids"].tolist()
    # ids in collection-file order, the dict maps id -> row (index_text.py:88,107)
    file_ids = [int(ln.split("\t")[0]) for ln in open(os.path.join(FIX, "collection.tsv"))]
    assert ids.tolist() == file_ids and meta["text_id_to_idx"][file_ids[17]] == 17
    qids, D, I = _parse_run(_golden_run(), 1000)
    assert qids.tolist() == [int(ln.split("\t")[0]) for ln in open(os.path.join(FIX, "queries.dev.tsv"))]
    assert (I[:, 700:] == -1).all() and (I[:, :700] >= 7_000_000).all()
    assert _golden_run().splitlines()[-1].endswith(b"\t-1\t1000\t-3.4028234663852886e+38")


def test_index_text_reproduces_the_reference_index_file_and_meta(cldrd_lib, tmp_path, monkeypatch):
    """Our retriever/index_text.py (CPU here: the encoder is plain PyTorch, the writer is host code) with the arguments
    the reference script was given: same file name, same bytes, same meta.pkl, same printed lines."""
    from retriever import index_text
    monkeypatch.setenv("CLDRD_LOADER_WORKERS", "0")
    monkeypatch.delenv("WORLD_SIZE", raising=False)
    exp = _experiment(tmp_path)
    model_dir = os.path.join(FIX, "tiny-distilbert")
    args = index_text.get_args(["--resume", str(exp / "models" / "checkpoint_120000.pth.tar"), "--model_name_or_path", model_dir,
                                "--tokenizer_name_or_path", model_dir, "--passages_path", os.path.join(FIX, "collection.tsv"),
                                "--index_dir", str(exp / "index") + "/"])
    out = io.StringIO()
    with contextlib.redirect_stdout(out):
        path = index_text.main(args)
    assert os.path.basename(path) == "checkpoint_120000.index"          # Path(resume).stem.split(".")[0] + ".index"
    ours, gold = open(path, "rb").read(), open(os.path.join(FIX, "checkpoint_120000.index"), "rb").read()
    assert len(ours) == len(gold) == 82 + 700 * 64 * 4 + 8 + 700 * 8
    assert ours[:82] == gold[:82] and ours[82 + 700 * 64 * 4:] == gold[82 + 700 * 64 * 4:]     # headers, count, id_map
    a = np.frombuffer(ours, dtype=np.float32, count=700 * 64, offset=82)
    b = np.frombuffer(gold, dtype=np.float32, count=700 * 64, offset=82)
    # the rows are an fp32 encoder forward on the CPU in both: bit-equal on the machine that made the fixture,
    # within fp32 noise on another CPU (different oneDNN kernels)
    assert np.allclose(a, b, rtol=1e-5, atol=1e-6), float(np.abs(a - b).max())
    m_ours = pickle.load(open(exp / "index" / "meta.pkl", "rb"))
    m_gold = pickle.load(open(os.path.join(FIX, "meta.pkl"), "rb"))
    assert sorted(m_ours) == sorted(m_gold) == ["text_id_to_idx", "text_ids"]
    assert m_ours["text_ids"].dtype == m_gold["text_ids"].dtype and m_ours["text_ids"].tolist() == m_gold["text_ids"].tolist()
    assert m_ours["text_id_to_idx"] == m_gold["text_id_to_idx"]
    assert list(m_ours["text_id_to_idx"]) == list(m_gold["text_id_to_idx"])     # insertion order too
    printed = [ln.replace(str(tmp_path), "<work>") for ln in out.getvalue().splitlines() if ln.startswith(KEEP)]
    assert printed == _ref_stdout("index_text")


def test_query_encoder_mirror_reproduces_the_reference_embeddings(cldrd_lib, tmp_path):
    """DualEncoder + load_checkpoint + SequenceDataset + get_embeddings_from_scratch == models/nway_dual_encoder.py,
    dataset/sequence_dataset.py, retrieval_utils.py:30-58 as the reference ran them (query tower, max_length 30)."""
    import torch
    from torch.utils.data import DataLoader
    from transformers import AutoTokenizer
    from cldrd.encoder import DualEncoder, SequenceDataset, load_checkpoint
    from cldrd.retrieval_utils import get_embeddings_from_scratch
    model_dir = os.path.join(FIX, "tiny-distilbert")
    model = DualEncoder(model_dir, share_weights=False)
    with contextlib.redirect_stdout(io.StringIO()):
        load_checkpoint(model, os.path.join(FIX, "checkpoint_120000.pth.tar"), True)
    ds = SequenceDataset.create_from_seqs_file(os.path.join(FIX, "queries.dev.tsv"), AutoTokenizer.from_pretrained(model_dir), 30, True)
    with contextlib.redirect_stdout(io.StringIO()):
        embs, ids = get_embeddings_from_scratch(model, DataLoader(ds, batch_size=512, collate_fn=ds.collate_fn), True, True)
    gold = np.load(os.path.join(FIX, "query_embs.npy"))
    assert embs.dtype == np.float32 and embs.shape == gold.shape == (24, 64)
    assert np.allclose(embs, gold, rtol=1e-5, atol=1e-6), float(np.abs(embs - gold).max())
    assert ids == _parse_run(_golden_run(), 1000)[0].tolist()
    # the two towers differ in the checkpoint: the passage tower must NOT reproduce the query embeddings
    with torch.no_grad():
        tok = ds.collate_fn([ds[i] for i in range(len(ds))])["seq"]
        assert not np.allclose(model.passage_embs(tok).numpy(), gold, atol=1e-3)


class _OracleIndex:
    """What the reference's run was searched with (the generator's faiss stand-in): oracle.search behind `.search`."""

    def __init__(self, path):
        self.xb, self.ids, _ = O.read_index(path)

    def search(self, x, k):
        return O.search(self.xb, self.ids, x, k)


def test_retrieve_loop_and_writer_reproduce_the_reference_run_file(cldrd_lib, tmp_path):
    """index_retrieve(batch=128) -> regroup -> writer, ours against the bytes the reference's loops wrote
    (retrieve_top_passages.py:88-109), padding hits and all; then the array form + streamed writer our CLI uses."""
    import cldrd
    from cldrd.retrieval_utils import index_retrieve, index_retrieve_arrays
    index = _OracleIndex(os.path.join(FIX, "checkpoint_120000.index"))
    xq = np.load(os.path.join(FIX, "query_embs.npy"))
    gold = _golden_run()
    qids = [int(ln.split("\t")[0]) for ln in open(os.path.join(FIX, "queries.dev.tsv"))]
    out = io.StringIO()
    with contextlib.redirect_stdout(out):
        nn_scores, nn_ids = index_retrieve(index, xq, 1000, batch=128)        # lists of lists, like the reference
    assert isinstance(nn_scores, list) and isinstance(nn_scores[0][0], float) and isinstance(nn_ids[0][0], int)
    assert out.getvalue().splitlines()[0] == "Query Num 24"
    a = tmp_path / "runs" / "dev.run"
    avg = cldrd.write_run_file(str(a), qids, np.array(nn_ids, dtype=np.int64), np.array(nn_scores, dtype=np.float32))
    assert a.read_bytes() == gold
    assert f"average ranks per query = {avg}" == _ref_stdout("retrieve_top_passages")[-1]
    with contextlib.redirect_stdout(io.StringIO()):
        D, I = index_retrieve_arrays(index, xq, 1000)
    b = tmp_path / "runs" / "dev.stream.run"
    st = cldrd.RunFileStream(str(b))
    for lo in range(0, 24, 7):
        st.put(np.asarray(qids[lo:lo + 7], dtype=np.int64), I[lo:lo + 7], D[lo:lo + 7])
    assert st.close() == avg and b.read_bytes() == gold


def test_transposed_script_passages_to_top_queries(cldrd_lib, tmp_path):
    """retriever/retrieve_top_queries.py as the reference ran it (k = 200, shared-weight encoder without a checkpoint,
    max_length 256, no path guards): our encoder mirror reproduces its passage embeddings, our retrieve loop + writer its
    `pid \\t qid \\t rank \\t score` file byte for byte, and the mirrored CLI is wired with the same roles and the
    same summary line."""
    import cldrd
    from torch.utils.data import DataLoader
    from transformers import AutoTokenizer
    from cldrd.encoder import DualEncoder, SequenceDataset
    from cldrd.retrieval_utils import get_embeddings_from_scratch, index_retrieve_arrays
    from retriever import retrieve_top_queries
    model_dir = os.path.join(FIX, "tiny-distilbert")
    model = DualEncoder(model_dir, share_weights=True)
    ds = SequenceDataset.create_from_seqs_file(os.path.join(FIX, "passages.small.tsv"), AutoTokenizer.from_pretrained(model_dir), 256, False)
    with contextlib.redirect_stdout(io.StringIO()):
        embs, pids = get_embeddings_from_scratch(model, DataLoader(ds, batch_size=512, collate_fn=ds.collate_fn), True, False)
    gold_embs = np.load(os.path.join(FIX, "passage_embs.npy"))
    assert embs.shape == gold_embs.shape == (40, 64) and np.allclose(embs, gold_embs, rtol=1e-5, atol=1e-6)
    gold = _golden_run("passages.run.gz")
    p_ref, D_ref, I_ref = _parse_run(gold, 200)
    assert pids == p_ref.tolist() == [int(ln.split("\t")[0]) for ln in open(os.path.join(FIX, "passages.small.tsv"))]
    with contextlib.redirect_stdout(io.StringIO()):
        D, I = index_retrieve_arrays(_OracleIndex(os.path.join(FIX, "checkpoint_120000.index")), gold_embs, 200)
    out = tmp_path / "passages.run"
    avg = cldrd.write_run_file(str(out), pids, I, D)
    assert out.read_bytes() == gold
    ref_lines = _ref_stdout("retrieve_top_queries")
    assert ref_lines[-2:] == ["# unique passages = 40", f"average ranks per query = {avg}"]
    # the mirrored CLI: same defaults, roles swapped, the reference's summary line, no path guards
    a = retrieve_top_queries.get_args(["--passages_path", "x.tsv", "--index_path", "y", "--output_path", "z"])
    assert a.top_k == 200 and a.max_length == 256 and a.share_weights and not a.is_parallel and a.queries_path == "x.tsv"
    src = open(retrieve_top_queries.__file__).read()
    assert 'header="# unique passages"' in src and "is_query_side=False" in src and "guards=False" in src


def test_reader_takes_the_reference_built_index_file(cldrd_lib):
    import cldrd
    index = cldrd.read_index(os.path.join(FIX, "checkpoint_120000.index"))
    meta = pickle.load(open(os.path.join(FIX, "meta.pkl"), "rb"))
    assert isinstance(index, cldrd.IndexIDMap) and index.ntotal == 700 and index.d == 64
    assert index.id_map.tolist() == meta["text_ids"].tolist()
    xb, _, _ = O.read_index(os.path.join(FIX, "checkpoint_120000.index"))
    assert np.array_equal(index._rows.materialize(), xb)


@pytest.mark.gpu
@pytest.mark.parametrize("scan", ["f16", "tf32", "simt"])
def test_gpu_search_on_the_reference_built_index_matches_the_reference_run(cldrd_lib, tmp_path, scan):
    """read_index -> convert_index_to_gpu -> index_retrieve(batch=128) -> writer on the B200, against the run file the
    reference's scripts wrote from the same index file and query embeddings: parity rule (scores 1e-5 relative, ids
    positional except near-ties), identical padding, identical line structure."""
    import cldrd
    from cldrd.retrieval_utils import convert_index_to_gpu, index_retrieve
    xq = np.load(os.path.join(FIX, "query_embs.npy"))
    qids_ref, D_ref, I_ref = _parse_run(_golden_run(), 1000)
    index = cldrd.read_index(os.path.join(FIX, "checkpoint_120000.index"))
    co = cldrd.GpuClonerOptions()
    co.scan = scan
    gpu = cldrd.index_cpu_to_gpu(cldrd.StandardGpuResources(), 0, index, co)
    with contextlib.redirect_stdout(io.StringIO()):
        nn_scores, nn_ids = index_retrieve(gpu, xq, 1000, batch=128)
    D, I = np.array(nn_scores, dtype=np.float32), np.array(nn_ids, dtype=np.int64)
    assert np.array_equal(I[:, 700:], I_ref[:, 700:]) and np.array_equal(D[:, 700:], D_ref[:, 700:])     # the padding
    xb, ids, _ = O.read_index(os.path.join(FIX, "checkpoint_120000.index"))
    r = O.compare_topk(D[:, :700], I[:, :700], D_ref[:, :700], I_ref[:, :700], *O.search(xb, ids, xq, 700, dtype=np.float64))
    assert r["ok"] and r["overlap"] == 1.0, r
    run = tmp_path / "runs" / "dev.run"
    cldrd.write_run_file(str(run), qids_ref, I, D)
    ours, gold = run.read_bytes().splitlines(), _golden_run().splitlines()
    assert len(ours) == len(gold) == 24000
    assert [ln.split(b"\t")[:1] + ln.split(b"\t")[2:3] for ln in ours] == [ln.split(b"\t")[:1] + ln.split(b"\t")[2:3] for ln in gold]
    gpu.close()
    # the default call of the reference (one GPU through convert_index_to_gpu) takes the same path
    gpu2 = convert_index_to_gpu(cldrd.read_index(os.path.join(FIX, "checkpoint_120000.index")), 0, False)
    D2, I2 = gpu2.search(xq, 1000)
    assert np.array_equal(I2[:, 700:], I_ref[:, 700:])
    assert O.compare_topk(D2[:, :700], I2[:, :700], D_ref[:, :700], I_ref[:, :700],
                          *O.search(xb, ids, xq, 700, dtype=np.float64))["ok"]
