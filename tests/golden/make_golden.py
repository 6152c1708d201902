"""Generates the committed golden fixtures from the CPU oracle (oracle/flat_ip.py).

PARITY UNPINNED: the reference's search arithmetic is faiss (absent here, un-pinned upstream)
and the reference holds no golden vectors of its own, so these are authored from the restated
semantics.  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import flat_ip as O  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    # (1) hand-assembled 2-vector, d=4 IxMp{IxFI} file: literal bytes live in tests/test_oracle.py;
    #     also written here for the C reader tests.
    xb = np.array([[1.0, 2.0, 3.0, 4.0], [-1.0, 0.5, 0.25, 8.0]], dtype=np.float32)
    ids = np.array([7, 2 ** 33 + 5], dtype=np.int64)
    with open(os.path.join(HERE, "tiny_ixmp.index"), "wb") as f:
        f.write(O.write_index_bytes(xb, ids))
    # (2) seeded 1000x64 index, 16 queries, k=10 and k=100, permuted ids
    xb = O.synth(1000, 64, 0)
    xq = O.synth(16, 64, 1)
    ids = O.synth_ids(1000, 7)
    D10, I10 = O.search(xb, ids, xq, 10)
    D100, I100 = O.search(xb, ids, xq, 100)
    np.savez_compressed(os.path.join(HERE, "seeded_1000x64.npz"), D10=D10, I10=I10, D100=D100, I100=I100)
    # (3) 20000x768 index (config-1 shape scaled to keep the fixture small), 8 queries, k=1000:
    #     only the results are stored; inputs are regenerated from the seeds.
    xb = O.synth(20000, 768, 0)
    xq = O.synth(8, 768, 1)
    D, R = O.search_rows(xb, xq, 1000)
    np.savez_compressed(os.path.join(HERE, "seeded_20000x768_k1000.npz"), D=D, R=R.astype(np.int32))
    # (4) run-file text for a fixed result block
    Dr = np.array([[103.856, 71.5, 0.1, -2.25e-5], [1e16, 3.0, 1.5e-7, -0.0]], dtype=np.float32)
    Ir = np.array([[5, 2 ** 33 + 5, 0, -1], [9, 8, 7, 6]], dtype=np.int64)
    O.write_run(os.path.join(HERE, "run_golden.tsv"), [1048585, 2], Ir, Dr)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
