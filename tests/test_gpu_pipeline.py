"""End-to-end pipeline on the GPU with a tiny random-init DistilBERT built offline:
index_text -> index file + meta.pkl -> retrieve_top_passages -> run file, checked against the oracle
run on the same embeddings (the reference's three-script flow, README.md:16-36)."""
import os
import pickle
import sys

import numpy as np
import pytest

from oracle import flat_ip as O

pytestmark = pytest.mark.gpu

WORDS = ["alpha", "beta", "gamma", "delta", "river", "stone", "cloud", "tensor", "query", "passage", "index", "score",
         "blue", "green", "fast", "slow", "north", "south", "model", "train", "dev", "rank", "deep", "dense"]


def _tiny_model_dir(tmp_path):
    from transformers import BertTokenizerFast, DistilBertConfig, DistilBertModel
    import torch
    d = tmp_path / "tiny-distilbert"
    d.mkdir()
    vocab = ["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"] + WORDS
    (d / "vocab.txt").write_text("\n".join(vocab) + "\n")
    tok = BertTokenizerFast(vocab_file=str(d / "vocab.txt"), do_lower_case=True)
    tok.save_pretrained(str(d))
    torch.manual_seed(2)
    cfg = DistilBertConfig(vocab_size=len(vocab), dim=64, n_layers=2, n_heads=4, hidden_dim=128, max_position_embeddings=64)
    DistilBertModel(cfg).save_pretrained(str(d))
    return str(d)


def test_index_text_then_retrieve_top_passages(cldrd_lib, tmp_path):
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "cl-drd_b200"))
    import torch
    from torch.utils.data import DataLoader
    from transformers import AutoTokenizer
    from cldrd.encoder import DualEncoder, SequenceDataset
    from cldrd.retrieval_utils import get_embeddings_from_scratch
    from retriever import index_text, retrieve_top_passages, retrieve_top_queries
    model_dir = _tiny_model_dir(tmp_path)
    rng = np.random.default_rng(0)
    coll, quer = tmp_path / "collection.tsv", tmp_path / "queries.dev.tsv"
    pids = rng.permutation(5000)[:300] + 7_000_000
    with open(coll, "w") as f:
        for pid in pids:
            f.write(f"{pid}\t{' '.join(rng.choice(WORDS, size=rng.integers(5, 30)))}\n")
    qids = rng.permutation(1000)[:25] + 1_048_000
    with open(quer, "w") as f:
        for qid in qids:
            f.write(f"{qid}\t{' '.join(rng.choice(WORDS, size=rng.integers(2, 8)))}\n")
    index_dir = str(tmp_path / "index") + "/"
    a = index_text.get_args(["--model_name_or_path", model_dir, "--tokenizer_name_or_path", model_dir, "--passages_path", str(coll),
                             "--index_dir", index_dir, "--index_name", "checkpoint_120000", "--share_weights", "--batch_size", "64"])
    torch.manual_seed(3)
    index_path = index_text.main(a)
    assert index_path.endswith("checkpoint_120000.index")
    xb, ids, info = O.read_index(index_path)                       # the oracle's reader parses our file
    assert info["fourcc"] == "IxMp" and xb.shape == (300, 64) and ids.tolist() == pids.tolist()
    meta = pickle.load(open(os.path.join(index_dir, "meta.pkl"), "rb"))
    assert meta["text_ids"].tolist() == pids.tolist() and meta["text_id_to_idx"][int(pids[5])] == 5
    run = tmp_path / "runs" / "dev.run"
    b = retrieve_top_passages.get_args(["--model_name_or_path", model_dir, "--tokenizer_name_or_path", model_dir,
                                        "--queries_path", str(quer), "--index_path", index_path, "--top_k", "20",
                                        "--output_path", str(run), "--share_weights"])
    retrieve_top_passages.main(b)
    lines = [ln.split("\t") for ln in run.read_text().splitlines()]
    assert len(lines) == 25 * 20 and [int(ln[2]) for ln in lines[:20]] == list(range(1, 21))
    assert [int(ln[0]) for ln in lines[::20]] == qids.tolist()   # query-file order
    # oracle on the same embeddings
    model = DualEncoder(model_dir, share_weights=True).cuda()
    tok = AutoTokenizer.from_pretrained(model_dir)
    ds = SequenceDataset.create_from_seqs_file(str(quer), tok, 30, is_query=True)
    xq, ids_q = get_embeddings_from_scratch(model, DataLoader(ds, batch_size=512, collate_fn=ds.collate_fn), True, True)
    assert ids_q == qids.tolist()
    D_ref, I_ref = O.search(xb, ids, xq, 20)
    D_run = np.array([float(ln[3]) for ln in lines], dtype=np.float32).reshape(25, 20)
    I_run = np.array([int(ln[1]) for ln in lines], dtype=np.int64).reshape(25, 20)
    r = O.compare_topk(D_run, I_run, D_ref, I_ref, *O.search(xb, ids, xq, 36, dtype=np.float64))
    assert r["ok"], r
    # streamed form (query sets above SEARCH_CHUNK: the run file grows chunk by chunk behind the search): same bytes
    run_s = tmp_path / "runs" / "dev.streamed.run"
    b.output_path = str(run_s)
    old_chunk, retrieve_top_passages.SEARCH_CHUNK = retrieve_top_passages.SEARCH_CHUNK, 7
    try:
        retrieve_top_passages.main(b)
    finally:
        retrieve_top_passages.SEARCH_CHUNK = old_chunk
    assert run_s.read_bytes() == run.read_bytes()
    # the transposed script: passages against the same index, top-5
    run2 = tmp_path / "runs" / "passages.run"
    c = retrieve_top_queries.get_args(["--model_name_or_path", model_dir, "--tokenizer_name_or_path", model_dir,
                                       "--passages_path", str(coll), "--index_path", index_path, "--top_k", "5",
                                       "--output_path", str(run2)])
    retrieve_top_queries.main(c)
    l2 = [ln.split("\t") for ln in run2.read_text().splitlines()]
    assert len(l2) == 300 * 5
    assert [int(ln[2]) for ln in l2[:5]] == [1, 2, 3, 4, 5]
    assert [int(ln[0]) for ln in l2[::5]] == pids.tolist()          # passage-file order
    s2 = np.array([float(ln[3]) for ln in l2]).reshape(300, 5)
    assert (np.diff(s2, axis=1) <= 0).all()
    assert set(int(ln[1]) for ln in l2) <= set(pids.tolist())


def test_retrieve_top_passages_under_torchrun_equals_single_process(cldrd_lib, tmp_path):
    """The CLI in its one-process-per-GPU form (needs 2 GPUs): same run file, byte for byte, as the single-process run."""
    import subprocess
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "cl-drd_b200"))
    from retriever import index_text, retrieve_top_passages
    model_dir = _tiny_model_dir(tmp_path)
    rng = np.random.default_rng(5)
    coll, quer = tmp_path / "collection.tsv", tmp_path / "queries.dev.tsv"
    pids = rng.permutation(5000)[:400] + 7_000_000
    with open(coll, "w") as f:
        for pid in pids:
            f.write(f"{pid}\t{' '.join(rng.choice(WORDS, size=rng.integers(5, 30)))}\n")
    qids = rng.permutation(1000)[:31] + 1_048_000
    with open(quer, "w") as f:
        for qid in qids:
            f.write(f"{qid}\t{' '.join(rng.choice(WORDS, size=rng.integers(2, 8)))}\n")
    index_dir = str(tmp_path / "index") + "/"
    a = index_text.get_args(["--model_name_or_path", model_dir, "--tokenizer_name_or_path", model_dir, "--passages_path", str(coll),
                             "--index_dir", index_dir, "--index_name", "checkpoint_1", "--share_weights", "--batch_size", "64"])
    index_path = index_text.main(a)
    common = ["--model_name_or_path", model_dir, "--tokenizer_name_or_path", model_dir, "--queries_path", str(quer),
              "--index_path", index_path, "--top_k", "20", "--share_weights"]
    run1 = tmp_path / "runs" / "dev.one.run"
    retrieve_top_passages.main(retrieve_top_passages.get_args(common + ["--output_path", str(run1)]))
    run2 = tmp_path / "runs" / "dev.two.run"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29537", os.path.join(root, "cl-drd_b200", "retriever", "retrieve_top_passages.py")] + common + \
          ["--output_path", str(run2)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-6000:]
    assert run2.read_bytes() == run1.read_bytes()



def test_index_text_under_torchrun_on_two_gpus_is_byte_identical(cldrd_lib, tmp_path):
    """The sharded index build on real GPUs (needs 2; NCCL group, fp16 autocast encoder on each rank's own device):
    rank r encodes and writes its row range of the one file; same bytes and meta.pkl as the single-process build."""
    import subprocess
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    model_dir = _tiny_model_dir(tmp_path)
    rng = np.random.default_rng(9)
    coll = tmp_path / "collection.tsv"
    pids = rng.permutation(100_000)[:512] + 7_000_000
    with open(coll, "w") as f:
        for pid in pids:
            f.write(f"{pid}\t{' '.join(rng.choice(WORDS, size=rng.integers(5, 30)))}\n")
    script = os.path.join(root, "cl-drd_b200", "retriever", "index_text.py")
    # 512 rows, batches of 128: rank r's batches are exactly batches 2r, 2r+1 of the single-process run (same padding)
    common = ["--model_name_or_path", model_dir, "--tokenizer_name_or_path", model_dir, "--passages_path", str(coll),
              "--index_name", "ckpt", "--share_weights", "--batch_size", "128", "--max_length", "40"]
    one, two = str(tmp_path / "one") + "/", str(tmp_path / "two") + "/"
    r = subprocess.run([sys.executable, script] + common + ["--index_dir", one], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29557", script] + common + ["--index_dir", two], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    a, b = open(os.path.join(one, "ckpt.index"), "rb").read(), open(os.path.join(two, "ckpt.index"), "rb").read()
    assert len(a) == len(b) == 82 + 512 * 64 * 4 + 8 + 512 * 8
    assert a == b
    xb, ids, info = O.read_index(os.path.join(two, "ckpt.index"))
    assert info["fourcc"] == "IxMp" and ids.tolist() == pids.tolist() and np.isfinite(xb).all() and np.abs(xb).sum() > 0
    ma, mb = (pickle.load(open(os.path.join(d, "meta.pkl"), "rb")) for d in (one, two))
    assert ma["text_ids"].tolist() == mb["text_ids"].tolist() == pids.tolist() and ma["text_id_to_idx"] == mb["text_id_to_idx"]
