"""Sharded index build (index_text.py under torchrun, world size 2, gloo, CPU): every rank encodes its slice of the
collection and writes it at its final offsets of the one index file; the result must be byte-identical to the
single-process build, meta.pkl included (retriever/index_text.py:86-109 is what both replace)."""
import contextlib
import io
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


def _build_in_process(argv, monkeypatch, **env):
    """index_text.main in this process (CPU; the subprocess form costs a fresh `import torch` per run)."""
    sys.path.insert(0, os.path.join(ROOT, "cl-drd_b200"))
    from retriever import index_text
    for k in ("WORLD_SIZE", "RANK", "LOCAL_RANK", "CLDRD_FAULT_BUILD_AFTER_ROWS"):
        monkeypatch.delenv(k, raising=False)
    monkeypatch.setenv("CLDRD_LOADER_WORKERS", "0")
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    out = io.StringIO()
    with contextlib.redirect_stdout(out):
        index_text.main(index_text.get_args(argv))
    return out.getvalue()


def test_sharded_build_is_byte_identical(cldrd_lib, tmp_path, monkeypatch):
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
    _build_in_process(common + ["--index_dir", one], monkeypatch)
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


def _collection(tmp_path, n=320):
    rng = np.random.default_rng(1)
    coll = tmp_path / "collection.tsv"
    pids = rng.permutation(100_000)[:n] + 7_000_000
    with open(coll, "w") as f:
        for pid in pids:
            f.write(f"{pid}\t{' '.join(rng.choice(WORDS, size=rng.integers(3, 20)))}\n")
    return coll, pids


def test_interrupted_build_continues_to_the_same_bytes(cldrd_lib, tmp_path, monkeypatch):
    """--continue_build: a build that died after 128 of 320 rows (injected failure) is completed from its progress
    record: same file and meta.pkl as an uninterrupted build, progress record gone; a record that does not belong to
    this build (other batch size) or a file of another shape starts over."""
    import json
    import pytest
    model_dir = _tiny_model_dir(tmp_path)
    coll, pids = _collection(tmp_path)
    common = ["--model_name_or_path", model_dir, "--tokenizer_name_or_path", model_dir, "--passages_path", str(coll),
              "--index_name", "ckpt", "--share_weights", "--batch_size", "64", "--max_length", "32"]
    whole, cont = str(tmp_path / "whole") + "/", str(tmp_path / "cont") + "/"

    def run(extra, **env):
        return _build_in_process(common + extra, monkeypatch, CLDRD_BUILD_SYNC_ROWS="64", **env)
    run(["--index_dir", whole])
    assert not [f for f in os.listdir(whole) if "progress" in f]
    with pytest.raises(RuntimeError, match="injected failure"):
        run(["--index_dir", cont], CLDRD_FAULT_BUILD_AFTER_ROWS="128")
    prog = json.load(open(os.path.join(cont, "ckpt.index.progress.0of1")))
    assert prog["done"] == 128 and prog["n"] == 320 and prog["batch_size"] == 64
    assert os.path.getsize(os.path.join(cont, "ckpt.index")) < os.path.getsize(os.path.join(whole, "ckpt.index"))
    out = run(["--index_dir", cont, "--continue_build"])
    assert "continuing the build at row 128" in out
    a, b = open(os.path.join(whole, "ckpt.index"), "rb").read(), open(os.path.join(cont, "ckpt.index"), "rb").read()
    assert a == b
    assert open(os.path.join(whole, "meta.pkl"), "rb").read() == open(os.path.join(cont, "meta.pkl"), "rb").read()
    assert not [f for f in os.listdir(cont) if "progress" in f]
    # a record of another batch size is not trusted: the build starts over (and still ends with the same bytes)
    with pytest.raises(RuntimeError, match="injected failure"):
        run(["--index_dir", cont], CLDRD_FAULT_BUILD_AFTER_ROWS="192")
    prog_path = os.path.join(cont, "ckpt.index.progress.0of1")
    prog = json.load(open(prog_path))
    prog["batch_size"] = 32
    json.dump(prog, open(prog_path, "w"))
    out = run(["--index_dir", cont, "--continue_build"])
    assert "continuing the build" not in out
    assert open(os.path.join(cont, "ckpt.index"), "rb").read() == a
    # --continue_build with nothing to continue, and with a file of another shape in the way: plain builds
    other = str(tmp_path / "other") + "/"
    run(["--index_dir", other, "--continue_build"])
    assert open(os.path.join(other, "ckpt.index"), "rb").read() == a
    O.write_index(os.path.join(other, "ckpt.index"), O.synth(50, 32, 0), O.synth_ids(50))
    run(["--index_dir", other, "--continue_build"])
    assert open(os.path.join(other, "ckpt.index"), "rb").read() == a


def test_interrupted_sharded_build_continues_to_the_same_bytes(cldrd_lib, tmp_path, monkeypatch):
    """The same under torchrun (world size 2, gloo): rank 1 dies after 64 of its 160 rows, torchrun takes rank 0 down
    with it; the continued run skips what each rank had recorded and ends with the single-process bytes."""
    model_dir = _tiny_model_dir(tmp_path)
    coll, pids = _collection(tmp_path)
    script = os.path.join(ROOT, "cl-drd_b200", "retriever", "index_text.py")
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", CLDRD_LOADER_WORKERS="0", OMP_NUM_THREADS="1", CLDRD_BUILD_SYNC_ROWS="32")
    # 320 rows, 2 ranks x 160 rows, batches of 32: rank r's batches are batches 5r .. 5r+4 of the single-process run
    common = ["--model_name_or_path", model_dir, "--tokenizer_name_or_path", model_dir, "--passages_path", str(coll),
              "--index_name", "ckpt", "--share_weights", "--batch_size", "32", "--max_length", "32"]
    one, two = str(tmp_path / "one") + "/", str(tmp_path / "two") + "/"
    _build_in_process(common + ["--index_dir", one], monkeypatch)
    launch = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1"]
    wrapper = tmp_path / "fail_on_rank1.py"
    wrapper.write_text("import os, runpy, sys\n"
                       "if os.environ.get('RANK') == '1':\n    os.environ['CLDRD_FAULT_BUILD_AFTER_ROWS'] = '64'\n"
                       f"sys.argv[0] = {script!r}\nrunpy.run_path({script!r}, run_name='__main__')\n")
    r = subprocess.run(launch + ["--master-port", "29548", str(wrapper)] + common + ["--index_dir", two], env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode != 0
    assert os.path.exists(os.path.join(two, "ckpt.index.progress.1of2"))
    r = subprocess.run(launch + ["--master-port", "29549", script] + common + ["--index_dir", two, "--continue_build"], env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "[rank 1] continuing the build at row 224" in r.stdout
    assert open(os.path.join(one, "ckpt.index"), "rb").read() == open(os.path.join(two, "ckpt.index"), "rb").read()
    assert open(os.path.join(one, "meta.pkl"), "rb").read() == open(os.path.join(two, "meta.pkl"), "rb").read()
    assert not [f for f in os.listdir(two) if "progress" in f]


def test_build_refuses_a_full_disk_before_encoding(cldrd_lib, tmp_path, monkeypatch):
    """The room for the whole index file is checked before the first batch is encoded (the reference finds out in
    faiss.write_index, after the 2.5 h of encoding); a continued build only needs what is still missing."""
    import pytest
    sys.path.insert(0, os.path.join(ROOT, "cl-drd_b200"))
    from retriever import index_text
    model_dir = _tiny_model_dir(tmp_path)
    coll, pids = _collection(tmp_path, n=64)
    common = ["--model_name_or_path", model_dir, "--tokenizer_name_or_path", model_dir, "--passages_path", str(coll),
              "--index_name", "ckpt", "--share_weights", "--batch_size", "32", "--max_length", "32"]
    need = 82 + 64 * 32 * 4 + 8 + 64 * 8
    monkeypatch.setattr(index_text, "_free_bytes", lambda d: need - 1)
    out = str(tmp_path / "full") + "/"
    with pytest.raises(OSError, match="not enough room"):
        _build_in_process(common + ["--index_dir", out], monkeypatch)
    assert not os.path.exists(os.path.join(out, "ckpt.index"))
    monkeypatch.setattr(index_text, "_free_bytes", lambda d: need)
    _build_in_process(common + ["--index_dir", out], monkeypatch)
    assert os.path.getsize(os.path.join(out, "ckpt.index")) == need
    # rebuilding over an existing file of the same size needs no extra room
    monkeypatch.setattr(index_text, "_free_bytes", lambda d: 0)
    _build_in_process(common + ["--index_dir", out], monkeypatch)
