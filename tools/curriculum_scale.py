"""Config 5's host-side post-processing at full scale, CPU only: a synthetic top-200 run of 502 939 queries
(100.6 M lines) written by the native writer, read back by the native reader, cut into curriculum groups.

    python tools/curriculum_scale.py [out.json]
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "cl-drd_b200"))
import cldrd  # noqa: E402
from cldrd import curriculum as CU  # noqa: E402


def main():
    nq, k = 502_939, 200
    rng = np.random.default_rng(0)
    I = rng.integers(0, 8_841_823, (nq, k), dtype=np.int64)
    D = np.sort(rng.standard_normal((nq, k), dtype=np.float32) * 3 + 100, axis=1)[:, ::-1].copy()
    qids = rng.permutation(1_200_000)[:nq].astype(np.int64)
    run = "/dev/shm/cldrd_scale.run"
    out = {"queries": nq, "k": k, "lines": nq * k, "host_cores": os.cpu_count()}
    try:
        t = time.perf_counter()
        cldrd.write_run_file(run, qids, I, D)
        out["write_s"] = round(time.perf_counter() - t, 2)
        out["file_bytes"] = os.path.getsize(run)
        t = time.perf_counter()
        q, lists = CU.read_run(run)
        out["read_s"] = round(time.perf_counter() - t, 2)
        assert np.array_equal(q, qids) and len(lists) == nq and np.array_equal(lists[12345], I[12345])
        t = time.perf_counter()
        ex = CU.groups_for_label_mode(q, lists, "9", seed=0, strict=False)
        out["groups_s"] = round(time.perf_counter() - t, 2)
        out["examples"] = len(ex)
        t = time.perf_counter()
        CU.write_groups("/dev/shm/cldrd_scale.groups.json", ex)
        out["write_groups_s"] = round(time.perf_counter() - t, 2)
    finally:
        for f in (run, "/dev/shm/cldrd_scale.groups.json"):
            if os.path.exists(f):
                os.unlink(f)
    out["write_lines_per_s"] = round(nq * k / out["write_s"])
    out["read_lines_per_s"] = round(nq * k / out["read_s"])
    print(json.dumps(out))
    if len(sys.argv) > 1:
        with open(sys.argv[1], "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
