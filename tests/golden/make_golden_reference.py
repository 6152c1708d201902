"""Golden fixtures produced by RUNNING THE REFERENCE'S OWN SCRIPTS in this container (CPU, no GPU, no faiss).

    python tests/golden/make_golden_reference.py          # needs /root/reference; writes tests/golden/ref_pipeline/

What runs unmodified, from /root/reference, through its own entry points:

    retriever/index_text.py            get_args() + main()     -> <index_dir>/checkpoint_120000.index, meta.pkl
    retriever/retrieve_top_passages.py get_args() + main()     -> dev.run
    retriever/retrieve_top_queries.py  get_args() + main()     -> passages.run   (k = 200, 40 passages against the same index)
    (and through them dataset/sequence_dataset.py, models/nway_dual_encoder.py,
     retriever/retrieval_utils.py: get_embeddings_from_scratch, convert_index_to_gpu, index_retrieve)

What is stubbed so that they run here, and nothing else:

    faiss                the module is absent (and un-pinned upstream): a stand-in whose IndexIDMap(IndexFlatIP) stores
                         the rows, whose write_index / read_index are oracle/flat_ip.py's and whose search IS
                         oracle.search.  So the arithmetic and the index-file BYTES in these fixtures are the oracle's
                         (PARITY UNPINNED for those two, as everywhere in this repo); everything around them -- TSV
                         parsing, tokenisation, the encoder loop, ids and their order, file naming, meta.pkl, the
                         query-batch loop, `.tolist()` boxing, the regroup dict, the f-string writer, the printed
                         summary lines -- is the reference's own code executing.
    .cuda()              Tensor.cuda / Module.cuda return self (no GPU here): the encoder runs in fp32 on the CPU;
                         torch.cuda.amp.autocast(enabled=True) is a no-op without CUDA.
    transformers.AdamW   removed from transformers 5.x; models/nway_dual_encoder.py:3 imports it (never uses it here).
    ujson                absent; dataset/nway_dataset.py:9 imports it (never used on this path) -> json.
    models.dual_encoder  retrieve_top_queries.py:23 imports `DualEncoder` from a module the reference does not ship: aliased
                         to the reference's own models/nway_dual_encoder.py:NwayDualEncoder (same constructor arguments,
                         same passage_embs).

Inputs are generated here from seeds and committed with the outputs (all small): a random-init DistilBERT-shaped
two-tower model (dim 64, 1 layer, 29-word vocabulary) as an HF model directory + a DataParallel-style checkpoint
(`module.` prefixes, passage tower re-initialised so the two towers differ), a 700-passage collection and 24 dev
queries.  k stays at the script's default 1000 > 700 rows, so every query also carries 300 padding hits
(id -1, score -FLT_MAX) through the reference's loops.

Consumers: tests/test_reference_pipeline.py (CPU: our index_text / encoder mirror / writer reproduce these files byte
for byte; GPU: our search on the fixture index against the fixture run under the parity rule).
"""
import gzip
import json
import os
import pickle
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
OUT = os.path.join(HERE, "ref_pipeline")

WORDS = ["alpha", "beta", "gamma", "delta", "river", "stone", "cloud", "tensor", "query", "passage", "index", "score",
         "blue", "green", "fast", "slow", "north", "south", "model", "train", "dev", "rank", "deep", "dense"]

# Runs inside the subprocess, cwd = /root/reference: installs the stubs, then drives the two scripts.
BOOT = r'''
import json, os, sys, types
sys.path[:0] = [%(ref)r, %(root)r]
import numpy as np, torch, transformers
if not hasattr(transformers, "AdamW"):
    transformers.AdamW = torch.optim.AdamW
sys.modules.setdefault("ujson", json)
torch.Tensor.cuda = lambda self, *a, **k: self
torch.nn.Module.cuda = lambda self, *a, **k: self
from oracle import flat_ip as O

faiss = types.ModuleType("faiss")
faiss.METRIC_INNER_PRODUCT = 0
SEARCH_LOG = []
class IndexFlatIP:
    def __init__(self, d): self.d, self.xb = int(d), np.empty((0, int(d)), np.float32)
    @property
    def ntotal(self): return self.xb.shape[0]
class IndexIDMap:
    def __init__(self, index): self.index, self.d, self.ids = index, index.d, np.empty((0,), np.int64)
    @property
    def ntotal(self): return self.index.ntotal
    def add_with_ids(self, x, ids):
        assert x.dtype == np.float32 and x.flags.c_contiguous and ids.dtype == np.int64
        self.index.xb = np.concatenate([self.index.xb, x]); self.ids = np.concatenate([self.ids, ids])
    def search(self, x, k):
        assert x.dtype == np.float32 and isinstance(k, int)
        SEARCH_LOG.append(np.array(x, copy=True))
        return O.search(self.index.xb, self.ids, x, k)
def write_index(index, path): O.write_index(path, index.index.xb, index.ids)
def read_index(path):
    xb, ids, info = O.read_index(path)
    assert info["fourcc"] == "IxMp"
    m = IndexIDMap(IndexFlatIP(xb.shape[1])); m.index.xb, m.ids = xb, ids
    return m
class StandardGpuResources:
    def setTempMemory(self, n): self.temp = n
class GpuClonerOptions:
    useFloat16 = False
def index_cpu_to_gpu(res, dev, index, co=None):
    assert dev == 0 and res.temp == 1024 * 1024 * 1024 and co.useFloat16 is False   # retrieval_utils.py:159-163
    return index
for n_ in ("IndexFlatIP", "IndexIDMap", "write_index", "read_index", "StandardGpuResources", "GpuClonerOptions",
           "index_cpu_to_gpu"):
    setattr(faiss, n_, globals()[n_])
sys.modules["faiss"] = faiss

work = %(work)r
import retriever.index_text as it
assert it.__file__.startswith(%(ref)r)
sys.argv = ["index_text.py", "--resume", work + "/experiment/models/checkpoint_120000.pth.tar",
            "--model_name_or_path", work + "/tiny-distilbert", "--tokenizer_name_or_path", work + "/tiny-distilbert",
            "--passages_path", work + "/collection.tsv", "--index_dir", work + "/experiment/index/"]
print("=== index_text ===")
it.main(it.get_args())
import retriever.retrieve_top_passages as rp
assert rp.__file__.startswith(%(ref)r)
sys.argv = ["retrieve_top_passages.py", "--resume", work + "/experiment/models/checkpoint_120000.pth.tar",
            "--model_name_or_path", work + "/tiny-distilbert", "--tokenizer_name_or_path", work + "/tiny-distilbert",
            "--queries_path", work + "/queries.dev.tsv", "--index_path", work + "/experiment/index/checkpoint_120000.index",
            "--output_path", work + "/runs/dev.run"]
print("=== retrieve_top_passages ===")
rp.main(rp.get_args())
np.save(work + "/query_embs.npy", np.concatenate(SEARCH_LOG))
del SEARCH_LOG[:]
import models.nway_dual_encoder as nde
alias = types.ModuleType("models.dual_encoder")
alias.DualEncoder = nde.NwayDualEncoder
sys.modules["models.dual_encoder"] = alias
import retriever.retrieve_top_queries as rq
assert rq.__file__.startswith(%(ref)r)
sys.argv = ["retrieve_top_queries.py", "--model_name_or_path", work + "/tiny-distilbert",
            "--tokenizer_name_or_path", work + "/tiny-distilbert", "--passages_path", work + "/passages.small.tsv",
            "--index_path", work + "/experiment/index/checkpoint_120000.index", "--output_path", work + "/runs/passages.run"]
print("=== retrieve_top_queries ===")
rq.main(rq.get_args())
np.save(work + "/passage_embs.npy", np.concatenate(SEARCH_LOG))
'''


def build_inputs(work):
    import torch
    from transformers import BertTokenizerFast, DistilBertConfig, DistilBertModel
    d = os.path.join(work, "tiny-distilbert")
    os.makedirs(d)
    vocab = ["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"] + WORDS
    with open(os.path.join(d, "vocab.txt"), "w") as f:
        f.write("\n".join(vocab) + "\n")
    BertTokenizerFast(vocab_file=os.path.join(d, "vocab.txt"), do_lower_case=True).save_pretrained(d)
    cfg = DistilBertConfig(vocab_size=len(vocab), dim=64, n_layers=1, n_heads=2, hidden_dim=64, max_position_embeddings=64)
    torch.manual_seed(2)
    q_tower = DistilBertModel(cfg)
    q_tower.save_pretrained(d)
    torch.manual_seed(12)
    p_tower = DistilBertModel(cfg)
    state = {}
    for k, v in q_tower.state_dict().items():
        state["module.query_encoder." + k] = v.clone()
    for k, v in p_tower.state_dict().items():
        state["module.passage_encoder." + k] = v.clone()
    os.makedirs(os.path.join(work, "experiment", "models"))
    torch.save({"state_dict": state}, os.path.join(work, "experiment", "models", "checkpoint_120000.pth.tar"))
    rng = np.random.default_rng(0)
    pids = rng.permutation(9000)[:700] + 7_000_000
    with open(os.path.join(work, "collection.tsv"), "w") as f:
        for pid in pids:
            f.write(f"{pid}\t{' '.join(rng.choice(WORDS, size=rng.integers(5, 30)))}\n")
    qids = rng.permutation(1000)[:24] + 1_048_000
    with open(os.path.join(work, "queries.dev.tsv"), "w") as f:
        for qid in qids:
            f.write(f"{qid}\t{' '.join(rng.choice(WORDS, size=rng.integers(2, 8)))}\n")
    with open(os.path.join(work, "passages.small.tsv"), "w") as f:
        for pid in rng.permutation(9000)[:40] + 8_000_000:
            f.write(f"{pid}\t{' '.join(rng.choice(WORDS, size=rng.integers(5, 30)))}\n")


def main():
    assert os.path.isdir(REF), "the reference tree is needed to make these fixtures"
    work = tempfile.mkdtemp(prefix="cldrd_ref_golden_")
    build_inputs(work)
    boot = BOOT % {"ref": REF, "root": ROOT, "work": work}
    r = subprocess.run([sys.executable, "-c", boot], cwd=REF, capture_output=True, text=True, timeout=1800)
    if r.returncode != 0:
        sys.stderr.write(r.stdout[-3000:] + "\n" + r.stderr[-6000:])
        raise SystemExit("the reference scripts failed")
    shutil.rmtree(OUT, ignore_errors=True)
    os.makedirs(OUT)
    for name in ("tiny-distilbert", "collection.tsv", "queries.dev.tsv", "query_embs.npy", "passages.small.tsv", "passage_embs.npy"):
        src = os.path.join(work, name)
        (shutil.copytree if os.path.isdir(src) else shutil.copy)(src, os.path.join(OUT, name))
    shutil.copy(os.path.join(work, "experiment", "models", "checkpoint_120000.pth.tar"), OUT)
    shutil.copy(os.path.join(work, "experiment", "index", "checkpoint_120000.index"), OUT)
    shutil.copy(os.path.join(work, "experiment", "index", "meta.pkl"), OUT)
    for run in ("dev.run", "passages.run"):
        with open(os.path.join(work, "runs", run), "rb") as f, gzip.GzipFile(os.path.join(OUT, run + ".gz"), "wb", mtime=0) as g:
            g.write(f.read())
    # the lines the scripts print that the mirrors must print too (progress bars and torch warnings dropped)
    keep = [ln for ln in r.stdout.splitlines()
            if ln.startswith(("===", "****", "load ", "# nan", "embs dtype", "retrieve ", "# unique", "average ranks", "Query Num"))]
    with open(os.path.join(OUT, "stdout.txt"), "w") as f:
        f.write("\n".join(ln.replace(work, "<work>") for ln in keep) + "\n")
    meta = pickle.load(open(os.path.join(OUT, "meta.pkl"), "rb"))
    info = {"made_by": "tests/golden/make_golden_reference.py", "reference_scripts_run": ["retriever/index_text.py", "retriever/retrieve_top_passages.py", "retriever/retrieve_top_queries.py"],
            "stubs": ["faiss (oracle-backed)", "Tensor.cuda/Module.cuda -> self", "transformers.AdamW", "ujson -> json",
                      "models.dual_encoder.DualEncoder -> models.nway_dual_encoder.NwayDualEncoder"],
            "passages": int(len(meta["text_ids"])), "queries": 24, "k": 1000, "dim": 64,
            "sizes": {n: os.path.getsize(os.path.join(OUT, n)) for n in sorted(os.listdir(OUT)) if os.path.isfile(os.path.join(OUT, n))}}
    with open(os.path.join(OUT, "README.json"), "w") as f:
        json.dump(info, f, indent=1)
    shutil.rmtree(work, ignore_errors=True)
    print(json.dumps(info))


if __name__ == "__main__":
    main()
