"""Where a 128-query search (the reference's loop shape) spends its time: device-only vs host API, per call."""
import os, sys, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cl-drd_b200")]
import torch
from cldrd import dist as CD

rows_n = int(os.environ.get("ROWS", 8_841_823))
reps = int(os.environ.get("REPS", 55))
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev).manual_seed(1000)
rows = torch.empty((rows_n, 768), dtype=torch.float32, device=dev)
for r0 in range(0, rows_n, 1 << 20):
    rows[r0:r0 + (1 << 20)].normal_(generator=g)
s = CD.ShardedSearcher.from_rows(rows, 0, rows_n, scan="f16")
del rows
q = torch.randn((reps * 128, 768), generator=g, dtype=torch.float32, device=dev)
q_np = q.cpu().numpy()
k = 1000
for i in range(4):
    s.local.search_device(q[i * 128:(i + 1) * 128], k)
    s.local.search(q_np[i * 128:(i + 1) * 128], k)
torch.cuda.synchronize()
out = {}
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter(); e0.record()
for i in range(reps):
    s.local.search_device(q[i * 128:(i + 1) * 128], k)
e1.record(); torch.cuda.synchronize()
out["device_loop_ms_per_call_events"] = e0.elapsed_time(e1) / reps
out["device_loop_ms_per_call_wall"] = (time.perf_counter() - t0) * 1e3 / reps
t0 = time.perf_counter()
for i in range(reps):
    s.local.search(q_np[i * 128:(i + 1) * 128], k)
out["host_loop_ms_per_call_wall"] = (time.perf_counter() - t0) * 1e3 / reps
s.shard.set_profiling(True)
sc = 0.0
t0 = time.perf_counter()
for i in range(reps):
    s.local.search_device(q[i * 128:(i + 1) * 128], k)
    sc += s.shard.scan_time()[0]
out["profiled_device_loop_ms_per_call_wall"] = (time.perf_counter() - t0) * 1e3 / reps
out["scan_ms_per_call"] = sc / reps
out["stats"] = s.shard.stats()
print(json.dumps(out))
