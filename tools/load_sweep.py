"""File -> HBM rate of cldrd_shard_load_file against the number of reader threads (CLDRD_LOAD_THREADS).
One GPU; the file lives in /dev/shm (page cache), like bench.py's index_load record.

    python tools/load_sweep.py [rows] > gpurun_out/load_sweep.json
"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cl-drd_b200"))
from cldrd import dist as CD  # noqa: E402
from cldrd._lib import check, lib, ptr  # noqa: E402

DIM = 768


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 21          # 6.4 GB
    path = f"/dev/shm/cldrd_load_sweep_{os.getpid()}.index"
    out = {"rows": n, "file_bytes": n * DIM * 4, "host_cores": os.cpu_count(), "runs": []}
    try:
        w = C.c_void_p()
        check(lib().cldrd_index_writer_begin(C.byref(w), path.encode(), n, DIM, 1, 0))
        block = np.random.Generator(np.random.PCG64(5)).standard_normal((1 << 15, DIM), dtype=np.float32)
        for r0 in range(0, n, 1 << 15):
            check(lib().cldrd_index_writer_append(w, ptr(block), min(1 << 15, n - r0)))
        ids = np.arange(n, dtype=np.int64)        # must outlive the call
        check(lib().cldrd_index_writer_finish(w, ptr(ids)))
        torch.zeros(1, device="cuda")
        for threads in (4, 1, 2, 4, 8, 12, 16, 24):
            os.environ["CLDRD_LOAD_THREADS"] = str(threads)
            best = 1e9
            for _ in range(2):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                s = CD.ShardedSearcher.from_file(path, 0, scan="f16")
                torch.cuda.synchronize()
                best = min(best, time.perf_counter() - t0)
                s.shard.close()
            out["runs"].append({"threads": threads, "seconds": round(best, 4), "gb_per_s": round(n * DIM * 4 / best / 1e9, 2)})
    finally:
        if os.path.exists(path):
            os.unlink(path)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
