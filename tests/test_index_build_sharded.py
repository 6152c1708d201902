"""Sharded index build (index_text.py under torchrun, world size 2, gloo, CPU): every rank encodes its slice of the
collection and writes it at its final offsets of the one index file; the result must be byte-identical to the
single-process build, meta.pkl included (retriever/index_text.py:86-109 is what both replace)."""
import os
import pickle
import subprocess
import sys

import numpy as np

from oracle import flat_ip as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORDS = ["alpha", "beta", "gamma", "delta", "river", "stone", "cloud", "tensor", "query", "passage", "index", "score"]


def _tiny_model_dir(tmp_path):
    from transformers import BertTokenizerFast, DistilBertConfig, DistilBertModel
    import torch
    d = tmp_path / "tiny-distilbert"
    d.mkdir()
    vocab = ["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"] + WORDS
    (d / "vocab.txt").write_text("\n".join(vocab) + "\n")
    BertTokenizerFast(vocab_file=str(d / "vocab.txt"), do_lower_case=True).save_pretrained(str(d))
    torch.manual_seed(2)
    DistilBertModel(DistilBertConfig(vocab_size=len(vocab), dim=32, n_layers=1, n_heads=2, hidden_dim=64,
                                     max_position_embeddings=64)).save_pretrained(str(d))
    return str(d)


def test_sharded_build_is_byte_identical(cldrd_lib, tmp_path):
    model_dir = _tiny_model_dir(tmp_path)
    rng = np.random.default_rng(0)
    coll = tmp_path / "collection.tsv"
    pids = rng.permutation(100_000)[:256] + 7_000_000
    with open(coll, "w") as f:
        for pid in pids:
            f.write(f"{pid}\t{' '.join(rng.choice(WORDS, size=rng.integers(3, 20)))}\n")
    script = os.path.join(ROOT, "cl-drd_b200", "retriever", "index_text.py")
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", CLDRD_LOADER_WORKERS="0", OMP_NUM_THREADS="1")
    # 256 rows, batches of 64: rank r's two batches are exactly batches 2r, 2r+1 of the single-process run (same padding)
    common = ["--model_name_or_path", model_dir, "--tokenizer_name_or_path", model_dir, "--passages_path", str(coll),
              "--index_name", "ckpt", "--share_weights", "--batch_size", "64", "--max_length", "32"]
    one, two = str(tmp_path / "one") + "/", str(tmp_path / "two") + "/"
    r = subprocess.run([sys.executable, script] + common + ["--index_dir", one], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29547", script] + common + ["--index_dir", two], env=env, capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    a, b = open(os.path.join(one, "ckpt.index"), "rb").read(), open(os.path.join(two, "ckpt.index"), "rb").read()
    assert len(a) == 82 + 256 * 32 * 4 + 8 + 256 * 8
    assert a == b
    xb, ids, info = O.read_index(os.path.join(two, "ckpt.index"))
    assert info["fourcc"] == "IxMp" and ids.tolist() == pids.tolist() and np.isfinite(xb).all() and np.abs(xb).sum() > 0
    ma, mb = (pickle.load(open(os.path.join(d, "meta.pkl"), "rb")) for d in (one, two))
    assert ma["text_ids"].tolist() == mb["text_ids"].tolist() == pids.tolist() and ma["text_id_to_idx"] == mb["text_id_to_idx"]


def test_ranged_writer_covers_disjoint_rows(cldrd_lib, tmp_path):
    """cldrd_index_writer_open_range: three writers, out of order, one file == cldrd_index_write."""
    import ctypes as C
    from cldrd._lib import check, ptr
    xb, ids = O.synth(1000, 24, 3), O.synth_ids(1000, 4)
    path = str(tmp_path / "r.index").encode()
    ws = []
    for row0, nrows, create in ((0, 300, 1), (300, 450, 0), (750, 250, 0)):
        w = C.c_void_p()
        check(cldrd_lib.cldrd_index_writer_open_range(C.byref(w), path, 1000, 24, 1, 0, row0, nrows, create))
        ws.append((w, row0, nrows))
    for w, row0, nrows in reversed(ws):                      # last range first
        part = np.ascontiguousarray(xb[row0:row0 + nrows])
        check(cldrd_lib.cldrd_index_writer_append(w, ptr(part[:100]), 100))
        check(cldrd_lib.cldrd_index_writer_append(w, ptr(part[100:]), nrows - 100))
    assert cldrd_lib.cldrd_index_writer_append(ws[1][0], ptr(xb), 1) != 0            # beyond its range
    for w, row0, _ in ws[1:]:
        check(cldrd_lib.cldrd_index_writer_finish(w, None))
    check(cldrd_lib.cldrd_index_writer_finish(ws[0][0], ptr(ids)))
    assert open(path, "rb").read() == O.write_index_bytes(xb, ids)
    w = C.c_void_p()
    assert cldrd_lib.cldrd_index_writer_open_range(C.byref(w), path, 1000, 24, 1, 0, 900, 200, 0) != 0   # outside the file
