"""Bring-up diagnostics for the scan kernels (run on the GPU box; writes gpurun_out/diag_<scan>.json).
Test infrastructure: compares raw scan scores and full searches against the CPU oracle and, on a
mismatch, prints enough structure (error by K block / row / column) to locate the bug."""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "cl-drd_b200")]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scan", default="f16")
    args = ap.parse_args()
    import torch
    import cldrd
    from cldrd._lib import check, lib
    from oracle import flat_ip as O
    out = {"scan": args.scan, "steps": []}

    def gpu_index(xb, scan):
        host = cldrd.IndexFlatIP(xb.shape[1])
        host.add(xb)
        co = cldrd.GpuClonerOptions()
        co.scan = scan
        return cldrd.index_cpu_to_gpu(cldrd.StandardGpuResources(), 0, host, co)

    try:
        for (N, d, nq, row_begin, nrows) in [(512, 64, 128, 0, 256), (512, 64, 128, 0, 512), (3000, 768, 200, 700, 1500),
                                             (9000, 768, 300, 256, 8192)]:
            xb, xq = O.synth(N, d, 10), O.synth(nq, d, 11)
            gpu = gpu_index(xb, args.scan)
            q = torch.from_numpy(xq).cuda()
            o = torch.full((nq, nrows), float("nan"), dtype=torch.float32, device="cuda")
            t0 = time.time()
            check(lib().cldrd_scan_dense_dev(gpu._shard.handle, C.c_void_p(q.data_ptr()), nq, row_begin, nrows,
                                             C.c_void_p(o.data_ptr()), None))
            torch.cuda.synchronize()
            got = o.cpu().numpy().astype(np.float64)
            ref = xq.astype(np.float64) @ xb[row_begin:row_begin + nrows].astype(np.float64).T
            scale = np.linalg.norm(xq, axis=1)[:, None] * np.linalg.norm(xb[row_begin:row_begin + nrows], axis=1)[None, :]
            err = np.abs(got - ref) / scale
            step = {"what": "dense", "shape": [N, d, nq, row_begin, nrows], "eff_scan": gpu.scan,
                    "max_rel_err": float(np.nanmax(err)), "nan": int(np.isnan(got).sum()),
                    "corr": float(np.corrcoef(np.nan_to_num(got).ravel(), ref.ravel())[0, 1]),
                    "secs": time.time() - t0}
            if step["max_rel_err"] > 1e-2 or step["nan"]:
                bad = np.nan_to_num(err, nan=1.0) > 1e-2
                step["bad_frac"] = float(bad.mean())
                step["bad_by_query_mod32"] = bad.reshape(nq, -1).mean(1)[:64].round(2).tolist()
                step["bad_by_col_first64"] = bad.mean(0)[:64].round(2).tolist()
                step["sample_got"] = got[:2, :8].tolist()
                step["sample_ref"] = ref[:2, :8].tolist()
                # does got match a partial-K product?  (locates descriptor-advance bugs)
                for kk in (8, 16, 32, 64, 128):
                    if kk <= d:
                        part = xq[:, :kk].astype(np.float64) @ xb[row_begin:row_begin + nrows, :kk].astype(np.float64).T
                        step[f"corr_partialK{kk}"] = float(np.corrcoef(np.nan_to_num(got).ravel(), part.ravel())[0, 1])
            out["steps"].append(step)
            print(json.dumps(step), flush=True)
            gpu.close()
        for (N, d, nq, k) in [(5000, 64, 33, 10), (30000, 128, 129, 100), (100000, 768, 256, 1000)]:
            xb, xq = O.synth(N, d, 20), O.synth(nq, d, 21)
            gpu = gpu_index(xb, args.scan)
            t0 = time.time()
            D, I = gpu.search(xq, k)
            secs = time.time() - t0
            D_ref, I_ref = O.search(xb, None, xq, k)
            D_ext, I_ext = O.search(xb, None, xq, k + 16, dtype=np.float64)
            r = O.compare_topk(D, I, D_ref, I_ref, D_ext, I_ext)
            step = {"what": "search", "shape": [N, d, nq, k], "cmp": r, "stats": gpu.last_stats(), "secs": secs}
            out["steps"].append(step)
            print(json.dumps(step), flush=True)
            gpu.close()
        out["ok"] = all((s.get("cmp", {}).get("ok", True) and s.get("max_rel_err", 0) < 1e-2) for s in out["steps"])
    except Exception as e:  # noqa: BLE001
        out["ok"] = False
        out["exception"] = repr(e)
        print("EXCEPTION", repr(e), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"diag_{args.scan}.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("DIAG", args.scan, "OK" if out.get("ok") else "FAILED")


if __name__ == "__main__":
    main()
