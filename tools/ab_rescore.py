"""A/B of the re-score kernel in the many-shard regime on ONE GPU: a 1/8 shard (1.1 M x 768), 6 980 queries, short
candidate lists (k = 150: ~200 entries per query, what a shard of an 8-GPU top-1000 search re-scores) and long ones
(k = 1000).  Run under ncu per library build:
    CLDRD_LIB_PATH=tools/ab/libcldrd_prev.so ncu --metrics gpu__time_duration.sum -k regex:rescore --csv python tools/ab_rescore.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cl-drd_b200")]
import torch  # noqa: E402
from cldrd import dist as CD  # noqa: E402

dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1000)
n = 8_841_823 // 8
rows = torch.empty((n, 768), dtype=torch.float32, device=dev)
for r0 in range(0, n, 1 << 20):
    rows[r0:r0 + (1 << 20)].normal_(generator=g)
q = torch.randn((6980, 768), generator=g, dtype=torch.float32, device=dev)
s = CD.ShardedSearcher.from_rows(rows, 0, n, scan="f16")
for k in (150, 1000):
    for _ in range(3):
        D, I = s.local.search_device(q, k, translate_ids=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30):
        D, I = s.local.search_device(q, k, translate_ids=False)
    e1.record()
    torch.cuda.synchronize()
    st = s.shard.stats()
    print(f"k={k} lib={os.environ.get('CLDRD_LIB_PATH', 'in-tree')} ms/search={e0.elapsed_time(e1) / 30:.3f} "
          f"rescored/q={st['rescored'] / 6980:.0f} max_list={st['max_list']}", flush=True)
