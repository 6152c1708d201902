"""Launch sequence of ONE rank's batch of the node-wide search, for ncu (which serialises kernels, so a multi-rank run
cannot be profiled: the flag barriers would wait for each other): a single-shard node over 1/8 of the index
(1 105 227 x 768), 6 980 queries, top-1000.  Every kernel of the protocol runs (prep, sample scan, sample top-j with the
peer-store path, barrier, levels + seed, filter scan, select, level counts, barrier, re-score with the counted cut and
key scatter, barrier, rank-placement merge + seed check + id map, barrier, status); with one shard the counted cut keeps
~k + band rows per query instead of the ~190 a rank of an 8-GPU search re-scores.
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_node.csv python tools/profile_node.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cl-drd_b200")]
import torch  # noqa: E402
from cldrd import dist as CD  # noqa: E402
from cldrd.index import GpuIndexShards  # noqa: E402

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1000)
n = 8_841_823 // 8
rows = torch.empty((n, 768), dtype=torch.float32, device=dev)
for r0 in range(0, n, 1 << 20):
    rows[r0:r0 + (1 << 20)].normal_(generator=g)
xq = torch.randn((6980, 768), generator=torch.Generator().manual_seed(1), dtype=torch.float32).pin_memory().numpy()
part = CD.ShardedSearcher.from_rows(rows, 0, n, scan="f16")
ids = np.random.Generator(np.random.PCG64(7)).permutation(n).astype(np.int64)
multi = GpuIndexShards([part.shard], ids, n, 768)
for _ in range(int(os.environ.get("REPS", "3"))):
    D, I = multi.search(xq, 1000)
print("node[world=1] stats", multi.last_stats(), "sorted", bool((np.diff(D, axis=1) <= 0).all()))
multi.close()
