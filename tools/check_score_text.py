"""Sweep of ALL 2^32 fp32 bit patterns through the run-file writer's score formatter (csrc/runfile.cpp:
shortest_f32, the specialised exact routine) against the general routine (std::to_chars digits laid out by
CPython's repr rules), plus a sample against Python's own repr(float(np.float32)).  CPU only, ~2-4 minutes on 8 cores.

    python tools/check_score_text.py [out.json]
"""
import ctypes as C
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "cl-drd_b200"))
from cldrd._lib import lib  # noqa: E402
import cldrd  # noqa: E402


def main():
    L = lib()
    chunk = 1 << 24

    def run(i):
        bad, fast = C.c_uint32(0), C.c_int64(0)
        n = L.cldrd_format_score_selfcheck(i * chunk, 1, chunk, C.byref(bad), C.byref(fast))
        return n, bad.value, fast.value

    t = time.time()
    with ThreadPoolExecutor(os.cpu_count()) as ex:       # ctypes releases the GIL
        res = list(ex.map(run, range(256)))
    differ = sum(r[0] for r in res)
    fast = sum(r[2] for r in res)
    first_bad = next((hex(r[1]) for r in res if r[0]), None)
    # Python's repr on a sample that covers every exponent, ties and powers of two
    rng = np.random.default_rng(0)
    bits = np.concatenate([
        rng.integers(0, 1 << 32, 300000, dtype=np.uint64).astype(np.uint32),
        (np.arange(256, dtype=np.uint32) << 23),                      # powers of two, zero, inf
        (np.arange(256, dtype=np.uint32) << 23) | np.uint32(1),
        (np.arange(256, dtype=np.uint32) << 23) | np.uint32(0x7fffff),
        (np.arange(256, dtype=np.uint32) << 23) | np.uint32(0x400000),
    ])
    vals = bits.view(np.float32)
    py_bad = 0
    for v in vals:
        if cldrd.format_score(v) != repr(float(v)):
            py_bad += 1
    out = {"patterns": 1 << 32, "differ_from_general_routine": int(differ), "first_bad": first_bad,
           "taken_by_fast_path": int(fast), "python_repr_sample": int(len(vals)), "python_repr_mismatches": py_bad,
           "seconds": round(time.time() - t, 1), "cores": os.cpu_count()}
    print(json.dumps(out))
    if len(sys.argv) > 1:
        with open(sys.argv[1], "w") as f:
            json.dump(out, f, indent=1)
    return 0 if differ == 0 and py_bad == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
