"""Small config under compute-sanitizer (memcheck / racecheck hooks):
    compute-sanitizer --tool memcheck  python tools/sanitize.py
    compute-sanitizer --tool racecheck python tools/sanitize.py
One single-shard search per scan mode (2-CTA tcgen05 kernel: remote st.shared::cluster + remote mbarrier arrives,
single-writer survivor segments parked in global memory between work units) and one node-wide search over two
in-process shards (peer stores, flag barriers), each checked against the oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cl-drd_b200")]


def main():
    import torch
    import cldrd
    from oracle import flat_ip as O
    n, d, nq, k = int(os.environ.get("SAN_ROWS", "6000")), 128, 260, 50
    xb, xq, ids = O.synth(n, d, 0), O.synth(nq, d, 1), O.synth_ids(n)
    D_ref, I_ref = O.search(xb, ids, xq, k)
    ext = O.search(xb, ids, xq, k + 16, dtype=np.float64)
    host = cldrd.IndexIDMap(cldrd.IndexFlatIP(d))
    host.add_with_ids(xb, ids)
    for scan in os.environ.get("SAN_SCANS", "f16,tf32,simt").split(","):
        co = cldrd.GpuClonerOptions()
        co.scan = scan
        gpu = cldrd.index_cpu_to_gpu(cldrd.StandardGpuResources(), 0, host, co)
        D, I = gpu.search(xq, k)
        r = O.compare_topk(D, I, D_ref, I_ref, *ext)
        print(f"single[{scan}] ok={r['ok']} stats={gpu.last_stats()}", flush=True)
        assert r["ok"], r
        gpu.close()
    co = cldrd.GpuMultipleClonerOptions()
    co.shard = True
    co.scan = "f16"
    multi = cldrd.index_cpu_to_gpu_multiple(None, [0, 0], host, co)
    D, I = multi.search(xq, k)
    r = O.compare_topk(D, I, D_ref, I_ref, *ext)
    print(f"node[2 shards] ok={r['ok']} stats={multi.last_stats()}", flush=True)
    assert r["ok"], r
    multi.close()
    torch.cuda.synchronize()
    print("sanitize run ok")


if __name__ == "__main__":
    main()
