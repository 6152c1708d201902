"""Small config under compute-sanitizer (memcheck / racecheck hooks):
    compute-sanitizer --tool memcheck  python tools/sanitize.py
    compute-sanitizer --tool racecheck python tools/sanitize.py
One single-shard search per scan mode (2-CTA tcgen05 kernel: remote st.shared::cluster + remote mbarrier arrives,
single-writer survivor segments parked in global memory between work units) and the node-wide protocol on a
single-shard node, unseeded and seeded, each checked against the oracle."""
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
    # The node-wide protocol (cldrd_node_*: peer stores, flag barriers, counted cut, key scatter, fused merge).  The
    # sanitizer serialises kernels of different streams, so two in-process shards on ONE GPU would wait for each other
    # until the barrier watchdog fires; a single-shard node (world = 1) runs every kernel of the protocol with the
    # barriers trivially open.  Small shard: unseeded batches; 2^20+ rows: sample -> levels -> counts -> cut.
    for rows_n, dd, seeded in ((n, d, False), (int(os.environ.get("SAN_BIG_ROWS", str((1 << 20) + 4096))), 64, True)):
        xb2, xq2 = (xb, xq) if not seeded else (O.synth(rows_n, dd, 2), O.synth(130, dd, 3))
        host2 = cldrd.IndexFlatIP(dd)
        host2.add(xb2)
        co = cldrd.GpuMultipleClonerOptions()
        co.shard = True
        co.scan = "f16"
        multi = cldrd.index_cpu_to_gpu_multiple(None, [0], host2, co)
        D, I = multi.search(xq2, k)
        D_r, I_r = O.search(xb2, None, xq2, k)
        r = O.compare_topk(D, I, D_r, I_r, *O.search(xb2, None, xq2, k + 16, dtype=np.float64))
        print(f"node[world=1, seeded={seeded}] ok={r['ok']} misses={multi.last_seed_misses} stats={multi.last_stats()}", flush=True)
        assert r["ok"], r
        multi.close()
    torch.cuda.synchronize()
    print("sanitize run ok")


if __name__ == "__main__":
    main()
