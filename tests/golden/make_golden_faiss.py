"""Pins the oracle to REAL faiss — to be run on any machine that has faiss (this container does not:
`import faiss` fails, no wheel on disk, no network; the reference pins no faiss version).

    python tests/golden/make_golden_faiss.py            # writes tests/golden/faiss_*.{index,npz}

For the seeded inputs the oracle's own fixtures use (PCG64 seeds 0 / 1 / 7, see oracle/flat_ip.py: synth, synth_ids)
it stores what the reference's calls produce:
  * faiss.write_index(IndexIDMap(IndexFlatIP(d)).add_with_ids(xb, ids))      retriever/index_text.py:91-105
      -> faiss_1000x64.index  (compared byte for byte with the oracle's and libcldrd's writers)
  * index.search(xq, k) on the CPU index                                     retriever/retrieval_utils.py:135
      -> faiss_1000x64.npz (k = 10, 100), faiss_20000x768_k1000.npz (k = 1000; rows as int32), faiss version inside
tests/test_oracle.py::test_faiss_made_fixtures consumes them when present and reports "parity unpinned" when absent.
Until such files are committed, every parity claim of this repo is against the restated semantics only.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))


def synth(n, d, seed):        # == oracle.flat_ip.synth (restated here so that this script needs nothing but numpy + faiss)
    return np.random.Generator(np.random.PCG64(seed)).standard_normal((n, d), dtype=np.float32)


def synth_ids(n, seed=7):     # == oracle.flat_ip.synth_ids
    return np.random.Generator(np.random.PCG64(seed)).permutation(n).astype(np.int64)


def main():
    import faiss
    ver = getattr(faiss, "__version__", "unknown")
    xb, xq, ids = synth(1000, 64, 0), synth(16, 64, 1), synth_ids(1000, 7)
    index = faiss.IndexIDMap(faiss.IndexFlatIP(64))
    index.add_with_ids(xb, ids)
    faiss.write_index(index, os.path.join(HERE, "faiss_1000x64.index"))
    D10, I10 = index.search(xq, 10)
    D100, I100 = index.search(xq, 100)
    np.savez_compressed(os.path.join(HERE, "faiss_1000x64.npz"), D10=D10, I10=I10, D100=D100, I100=I100, faiss_version=ver)
    xb, xq = synth(20000, 768, 0), synth(8, 768, 1)
    flat = faiss.IndexFlatIP(768)
    flat.add(xb)
    D, R = flat.search(xq, 1000)
    np.savez_compressed(os.path.join(HERE, "faiss_20000x768_k1000.npz"), D=D, R=R.astype(np.int32), faiss_version=ver)
    # fewer rows than k: faiss pads with id -1 / score -FLT_MAX (from faiss knowledge: pin it)
    small = faiss.IndexFlatIP(8)
    small.add(synth(5, 8, 2))
    Dp, Ip = small.search(synth(3, 8, 3), 8)
    np.savez_compressed(os.path.join(HERE, "faiss_padding.npz"), D=Dp, I=Ip, faiss_version=ver)
    print("faiss", ver, "fixtures written to", HERE)


if __name__ == "__main__":
    main()
