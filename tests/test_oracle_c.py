"""The C restatement of the oracle agrees with the numpy oracle (both are test infrastructure)."""
import ctypes as C
import os

import numpy as np

from oracle import flat_ip as O

SO = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "liboracle_flat_ip.so")


def _c_search(xb, ids, xq, k):
    if not os.path.exists(SO):
        import subprocess
        subprocess.run(["make", "-C", os.path.dirname(SO)], check=True)
    lib = C.CDLL(SO)
    lib.oracle_flat_ip_search.restype = C.c_int
    lib.oracle_flat_ip_search.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_int64, C.c_int,
                                          C.c_void_p, C.c_void_p]
    n, d = xb.shape
    nq = xq.shape[0]
    D = np.empty((nq, k), dtype=np.float32)
    I = np.empty((nq, k), dtype=np.int64)
    rc = lib.oracle_flat_ip_search(xb.ctypes.data, ids.ctypes.data if ids is not None else None, n, d,
                                   xq.ctypes.data, nq, k, D.ctypes.data, I.ctypes.data)
    assert rc == 0
    return D, I


def test_c_oracle_matches_numpy_oracle():
    xb, xq, ids = O.synth(3000, 96, 0), O.synth(9, 96, 1), O.synth_ids(3000)
    Dn, In = O.search(xb, ids, xq, 40)
    De, Ie = O.search(xb, ids, xq, 48, dtype=np.float64)
    Dc, Ic = _c_search(xb, ids, xq, 40)
    r = O.compare_topk(Dc, Ic, Dn, In, De, Ie)
    assert r["ok"], r


def test_c_oracle_ties_and_padding():
    xb = np.ones((5, 4), dtype=np.float32)
    xb[3] = 2.0
    q = np.ones((1, 4), dtype=np.float32)
    D, I = _c_search(xb, None, q, 8)
    assert I[0].tolist() == [3, 0, 1, 2, 4, -1, -1, -1]
    assert D[0, 5] == O.NEG_FLT_MAX
