"""GPU parity tests (B200): the CUDA path, called through the C ABI, against the CPU oracle.

Tolerance (BASELINE.json north_star): fp32 scores within 1e-5 relative; ids bit-exact except
inside near-tie bands of that width (oracle.compare_topk).  Integer outputs (ids, padding) are
compared exactly.
"""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import flat_ip as O

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")
SCANS = ["simt", "tf32", "f16", "bf16"]


def _gpu_index(xb, ids, scan, device=0):
    import cldrd
    host = cldrd.IndexIDMap(cldrd.IndexFlatIP(xb.shape[1])) if ids is not None else cldrd.IndexFlatIP(xb.shape[1])
    if ids is not None:
        host.add_with_ids(xb, ids)
    else:
        host.add(xb)
    co = cldrd.GpuClonerOptions()
    co.scan = scan
    return cldrd.index_cpu_to_gpu(cldrd.StandardGpuResources(), device, host, co)


def _check(gpu, xb, ids, xq, k, margin=16):
    D, I = gpu.search(xq, k)
    D_ref, I_ref = O.search(xb, ids, xq, k)
    D_ext, I_ext = O.search(xb, ids, xq, k + margin, dtype=np.float64)
    r = O.compare_topk(D, I, D_ref, I_ref, D_ext, I_ext)
    assert r["ok"], (gpu.scan, k, r, gpu.last_stats())
    assert r["overlap"] == 1.0
    # sortedness: descending scores within every row
    assert (np.diff(D, axis=1) <= 0).all()   # padding is -FLT_MAX, so this covers it too
    return D, I


# ------------------------------------------------------------------------------------------
# the scan kernels alone: raw scan scores against the fp64 product, inside the proven band
# ------------------------------------------------------------------------------------------

@pytest.mark.parametrize("scan", SCANS)
def test_scan_dense_within_error_bound(cldrd_lib, scan):
    import torch
    import cldrd
    from cldrd._lib import check
    xb, xq = O.synth(3000, 768, 10), O.synth(200, 768, 11)
    gpu = _gpu_index(xb, None, scan)
    assert gpu.scan == scan
    q = torch.from_numpy(xq).cuda()
    nrows, row_begin = 1500, 700   # unaligned begin, partial last tile
    out = torch.empty((xq.shape[0], nrows), dtype=torch.float32, device="cuda")
    check(cldrd_lib.cldrd_scan_dense_dev(gpu._shard.handle, C.c_void_p(q.data_ptr()), xq.shape[0], row_begin, nrows,
                                         C.c_void_p(out.data_ptr()), None))
    torch.cuda.synchronize()
    got = out.cpu().numpy().astype(np.float64)
    ref = xq.astype(np.float64) @ xb[row_begin:row_begin + nrows].astype(np.float64).T
    scale = np.linalg.norm(xq, axis=1)[:, None] * np.linalg.norm(xb[row_begin:row_begin + nrows], axis=1)[None, :]
    rel = np.abs(got - ref) / scale
    coef = {"simt": 2.0 ** -24 * 784, "tf32": 2.0 ** -9, "f16": 2.0 ** -10, "bf16": 2.0 ** -7}[scan]
    assert rel.max() <= coef, (scan, rel.max(), coef)
    # and it is a real product, not zeros: typical error far below the bound, result correlated
    assert np.corrcoef(got.ravel(), ref.ravel())[0, 1] > 0.9999
    gpu.close()


# ------------------------------------------------------------------------------------------
# search parity
# ------------------------------------------------------------------------------------------

@pytest.mark.parametrize("scan", SCANS)
def test_golden_seeded_1000x64(cldrd_lib, scan):
    g = np.load(os.path.join(GOLD, "seeded_1000x64.npz"))
    xb, xq, ids = O.synth(1000, 64, 0), O.synth(16, 64, 1), O.synth_ids(1000, 7)
    gpu = _gpu_index(xb, ids, scan)
    for k in (10, 100):
        D, I = gpu.search(xq, k)
        r = O.compare_topk(D, I, g[f"D{k}"], g[f"I{k}"])
        assert r["ok"], (scan, k, r)
    gpu.close()


@pytest.mark.parametrize("scan", SCANS)
def test_golden_20000x768_k1000(cldrd_lib, scan):
    g = np.load(os.path.join(GOLD, "seeded_20000x768_k1000.npz"))
    xb, xq = O.synth(20000, 768, 0), O.synth(8, 768, 1)
    gpu = _gpu_index(xb, None, scan)
    D, I = gpu.search(xq, 1000)
    D_ext, I_ext = O.search_rows(xb, xq, 1016, dtype=np.float64)
    r = O.compare_topk(D, I, g["D"], g["R"].astype(np.int64), D_ext, I_ext)
    assert r["ok"], (scan, r, gpu.last_stats())
    gpu.close()


@pytest.mark.parametrize("scan", SCANS)
@pytest.mark.parametrize("k", [1, 10, 1000, 2048])
def test_parity_medium(cldrd_lib, scan, k):
    xb, xq, ids = O.synth(30000, 128, 20), O.synth(129, 128, 21), O.synth_ids(30000, 22)
    gpu = _gpu_index(xb, ids, scan)
    _check(gpu, xb, ids, xq, k)
    st = gpu.last_stats()
    assert st["launches"] > 0 and st["chunks"] >= 2
    gpu.close()


@pytest.mark.parametrize("scan", ["f16", "tf32"])
def test_parity_config1_shape(cldrd_lib, scan):
    """BASELINE.json configs[0] shape (100k x 768, k=1000) on a query subset the oracle finishes fast."""
    xb, xq, ids = O.synth(100_000, 768, 0), O.synth(1000, 768, 1)[:200], O.synth_ids(100_000, 7)
    gpu = _gpu_index(xb, ids, scan)
    _check(gpu, xb, ids, xq, 1000)
    st = gpu.last_stats()
    assert st["tc_tiles"] > 0, "tensor-core scan did not run"
    assert st["fallback_queries"] == 0
    gpu.close()


@pytest.mark.parametrize("scan", SCANS)
@pytest.mark.parametrize("nq", [1, 127, 128, 129, 300])
def test_query_batch_edges(cldrd_lib, scan, nq):
    xb, xq = O.synth(9000, 64, 30), O.synth(nq, 64, 31)
    gpu = _gpu_index(xb, None, scan)
    _check(gpu, xb, None, xq, 20)
    gpu.close()


@pytest.mark.parametrize("scan", SCANS)
def test_fewer_rows_than_k_pads(cldrd_lib, scan):
    xb, xq, ids = O.synth(37, 32, 40), O.synth(5, 32, 41), O.synth_ids(37, 42) + 2 ** 33
    gpu = _gpu_index(xb, ids, scan)
    D, I = _check(gpu, xb, ids, xq, 100)
    assert (I[:, 37:] == -1).all() and (D[:, 37:] == O.NEG_FLT_MAX).all()
    assert (I[:, :37] >= 2 ** 33).all()  # ids beyond 2^31 survive
    gpu.close()


@pytest.mark.parametrize("scan", SCANS)
def test_exact_duplicates_tie_order(cldrd_lib, scan):
    """Duplicate rows: equal scores must come back lower-row-first; a flood of ties bigger than
    the candidate list forces the in-kernel exact compaction."""
    base = O.synth(50, 64, 50)
    xb = np.concatenate([base] * 200)          # 10000 rows, every vector 200 times
    xq = O.synth(7, 64, 51)
    gpu = _gpu_index(xb, None, scan)
    D, I = gpu.search(xq, 500)
    D_ref, I_ref = O.search(xb, None, xq, 500)
    np.testing.assert_allclose(D, D_ref, rtol=1e-5)
    # within a group of bit-equal scores rows are ascending
    for i in range(xq.shape[0]):
        same = D[i, 1:] == D[i, :-1]
        assert (I[i, 1:][same] > I[i, :-1][same]).all()
    r = O.compare_topk(D, I, D_ref, I_ref, *O.search(xb, None, xq, 700, dtype=np.float64))
    assert r["ok"], (scan, r, gpu.last_stats())
    gpu.close()
    # all rows identical: 6000 tied candidates > list capacity
    xb = np.tile(O.synth(1, 64, 52), (6000, 1))
    gpu = _gpu_index(xb, None, scan)
    D, I = gpu.search(xq, 1000)
    assert (I == np.arange(1000)[None, :]).all(), gpu.last_stats()
    gpu.close()


@pytest.mark.parametrize("scan", SCANS)
def test_negative_scores_only(cldrd_lib, scan):
    xb = np.abs(O.synth(5000, 64, 60))
    xq = -np.abs(O.synth(9, 64, 61))
    gpu = _gpu_index(xb, None, scan)
    D, I = _check(gpu, xb, None, xq, 50)
    assert (D < 0).all()
    gpu.close()


@pytest.mark.parametrize("scan", ["f16", "tf32", "simt"])
def test_adversarial_order_uses_dense_fallback(cldrd_lib, scan):
    """Rows sorted by ascending score for the queries: every chunk beats the running threshold,
    the survivor buffers overflow and the dense fallback must still return the exact answer."""
    d = 64
    qdir = O.synth(1, d, 70)[0]
    qdir /= np.linalg.norm(qdir)
    xb = O.synth(120_000, d, 71) * 0.05 + np.linspace(-3, 3, 120_000, dtype=np.float32)[:, None] * qdir[None, :]
    xb = xb.astype(np.float32)
    xq = (qdir[None, :] * 5 + O.synth(40, d, 72) * 0.01).astype(np.float32)
    gpu = _gpu_index(xb, None, scan)
    _check(gpu, xb, None, xq, 100)
    assert gpu.last_stats()["fallback_queries"] > 0, gpu.last_stats()
    gpu.close()


def test_unaligned_dim_falls_back_to_simt_scan(cldrd_lib):
    xb, xq = O.synth(4000, 30, 80), O.synth(11, 30, 81)
    gpu = _gpu_index(xb, None, "f16")
    assert gpu.scan == "simt"      # d*2 bytes is not a multiple of 16: TMA cannot describe it
    _check(gpu, xb, None, xq, 10)
    gpu.close()


def test_fp16_range_guard_and_auto_mode(cldrd_lib):
    import cldrd
    xb, xq = O.synth(3000, 64, 90) * 1e5, O.synth(4, 64, 91)
    with pytest.raises(cldrd.CldrdError):
        _gpu_index(xb, None, "f16")
    gpu = _gpu_index(xb, None, "auto")
    assert gpu.scan == "tf32"
    _check(gpu, xb, None, xq, 10)
    gpu.close()


def test_modes_agree_bit_exactly(cldrd_lib):
    """Every returned score comes from the same fp32 re-score routine: scan modes must agree to the bit."""
    xb, xq = O.synth(50_000, 256, 100), O.synth(64, 256, 101)
    res = {}
    for scan in SCANS:
        gpu = _gpu_index(xb, None, scan)
        res[scan] = gpu.search(xq, 100)
        gpu.close()
    for scan in SCANS[1:]:
        assert np.array_equal(res[scan][0], res["simt"][0]), scan
        assert np.array_equal(res[scan][1], res["simt"][1]), scan


def test_k_limits_and_errors(cldrd_lib):
    import cldrd
    xb = O.synth(100, 16, 110)
    gpu = _gpu_index(xb, None, "simt")
    with pytest.raises(RuntimeError):
        gpu.search(O.synth(1, 16, 0), 4096)
    with pytest.raises(TypeError):
        gpu.search(O.synth(1, 16, 0).astype(np.float64), 5)
    with pytest.raises(AssertionError):
        gpu.search(O.synth(1, 8, 0), 5)
    D, I = gpu.search(np.empty((0, 16), dtype=np.float32), 5)
    assert D.shape == (0, 5) and I.shape == (0, 5)
    gpu.close()


# ------------------------------------------------------------------------------------------
# file -> HBM -> search -> run file, the reference's own helpers driving our objects
# ------------------------------------------------------------------------------------------

def test_file_to_run_file_end_to_end(cldrd_lib, tmp_path):
    import cldrd
    from cldrd import retrieval_utils as RU
    xb, xq, ids = O.synth(20000, 96, 120), O.synth(300, 96, 121), O.synth_ids(20000, 122)
    path = tmp_path / "checkpoint_120000.index"
    O.write_index(str(path), xb, ids)                       # a file as real faiss would write it
    index = cldrd.read_index(str(path))
    index = RU.convert_index_to_gpu(index, 0, False)
    nn_scores, nn_ids = RU.index_retrieve(index, xq, 100, batch=128)   # reference loop shape
    assert isinstance(nn_scores, list) and len(nn_scores) == 300 and len(nn_ids[0]) == 100
    D_ref, I_ref = O.search(xb, ids, xq, 100)
    r = O.compare_topk(np.array(nn_scores, dtype=np.float32), np.array(nn_ids), D_ref, I_ref,
                       *O.search(xb, ids, xq, 116, dtype=np.float64))
    assert r["ok"], r
    D, I = RU.index_retrieve_arrays(index, xq, 100)
    assert np.array_equal(I, np.array(nn_ids))              # one pass == 128-query round trips
    qids = np.arange(1000, 1300, dtype=np.int64)
    run_a, run_b = tmp_path / "dev_a.run", tmp_path / "dev_b.run"
    cldrd.write_run_file(str(run_a), qids, I, D)
    O.write_run(str(run_b), qids.tolist(), nn_ids, nn_scores)
    assert run_a.read_bytes() == run_b.read_bytes()


def test_two_shards_merge_equals_single_shard(cldrd_lib):
    """Row-sharded search + merge kernel == single-shard search, bit for bit."""
    import torch
    from cldrd import dist as CD
    xb, xq, ids = O.synth(40_001, 128, 130), O.synth(150, 128, 131), O.synth_ids(40_001, 132)
    rows = torch.from_numpy(xb).cuda()
    q = torch.from_numpy(xq).cuda()
    id_map = torch.from_numpy(ids).cuda()
    one = CD.ShardedSearcher.from_rows(rows, 0, xb.shape[0], scan="f16", id_map=id_map)
    D1, I1 = one.search(q, 100)
    from cldrd.index import shard_ranges
    Ds, Is = [], []
    for rr in shard_ranges(xb.shape[0], 3):
        part = rows[rr.start:rr.stop].contiguous()
        sh = CD.ShardedSearcher.from_rows(part, rr.start, xb.shape[0], scan="f16")
        D, I = sh.local.search_device(q, 100, translate_ids=False)
        assert int(I.min()) >= rr.start and int(I.max()) < rr.stop
        Ds.append(D)
        Is.append(I)
    Dm, Im = CD.merge_candidates(torch.stack(Ds), torch.stack(Is), id_map)
    assert torch.equal(Dm, D1) and torch.equal(Im, I1)
    D_ref, I_ref = O.search(xb, ids, xq, 100)
    r = O.compare_topk(Dm.cpu().numpy(), Im.cpu().numpy(), D_ref, I_ref, *O.search(xb, ids, xq, 116, dtype=np.float64))
    assert r["ok"], r


def test_multi_gpu_in_process_shards(cldrd_lib):
    import torch
    import cldrd
    from cldrd import retrieval_utils as RU
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    xb, xq, ids = O.synth(30_000, 64, 140), O.synth(50, 64, 141), O.synth_ids(30_000, 142)
    host = cldrd.IndexIDMap(cldrd.IndexFlatIP(64))
    host.add_with_ids(xb, ids)
    gpu = RU.convert_index_to_gpu(host, [0, 1], False)
    D, I = gpu.search(xq, 50)
    D_ref, I_ref = O.search(xb, ids, xq, 50)
    r = O.compare_topk(D, I, D_ref, I_ref, *O.search(xb, ids, xq, 66, dtype=np.float64))
    assert r["ok"], r


# ------------------------------------------------------------------------------------------
# full size (BASELINE.json configs[1] shape): planted neighbours, sortedness, bit-equality of two scan precisions,
# and the oracle's comparison rule against a brute-force fp64 search on 64 queries.
# ------------------------------------------------------------------------------------------

def test_full_size_properties(cldrd_lib):
    import torch
    from cldrd import dist as CD
    free, _ = torch.cuda.mem_get_info()
    N, d, k = 8_841_823, 768, 1000
    if free < 60 * 2 ** 30:
        pytest.skip("needs ~45 GB of free HBM")
    g = torch.Generator(device="cuda").manual_seed(0)
    rows = torch.empty((N, d), dtype=torch.float32, device="cuda")
    step = 1 << 20
    for r0 in range(0, N, step):
        rows[r0:r0 + step].normal_(generator=g)
    nq = 256
    q = torch.randn((nq, d), generator=g, device="cuda")
    # planted neighbours: query i has 3 known rows with a dominant score
    planted = torch.randint(0, N, (nq, 3), generator=g, device="cuda")
    for j in range(3):
        rows[planted[:, j]] = q * (1.5 - 0.1 * j)
    s16 = CD.ShardedSearcher.from_rows(rows, 0, N, scan="f16")
    D16, I16 = s16.search(q, k)
    st = s16.shard.stats()
    assert st["tc_tiles"] > 0 and st["fallback_queries"] == 0, st
    assert torch.equal(I16[:, :3], planted), "planted neighbours not on top"
    assert (D16[:, 1:] <= D16[:, :-1]).all()
    # oracle on a slice: exact fp32 re-score of the returned rows reproduces the returned scores
    sel = I16[:8]
    ref = (rows[sel.reshape(-1)].double().reshape(8, k, d) * q[:8].double()[:, None, :]).sum(-1)
    assert torch.allclose(D16[:8].double(), ref, rtol=1e-5, atol=0)
    # the tf32 scan over the raw fp32 rows returns the same bits
    del s16
    s32 = CD.ShardedSearcher.from_rows(rows, 0, N, scan="tf32")
    D32, I32 = s32.search(q, k)
    assert torch.equal(D32, D16) and torch.equal(I32, I16)
    # the oracle's rule at full size: 64 queries against a brute-force fp64 search of all 8.8M rows (plain torch
    # matmul, chunked; bench.py's checker) through compare_topk -- scores within 1e-5 relative, ids exact outside
    # near-tie bands, overlap@1000 = 1.0
    import bench
    sel = torch.linspace(0, nq - 1, 64, device="cuda").long()
    s64, r64 = bench.brute_force_shard(torch, rows, 0, q[sel], k + 16)
    s64, r64 = s64.cpu().numpy(), r64.cpu().numpy()
    r = O.compare_topk(D16[sel].cpu().numpy(), I16[sel].cpu().numpy(), s64[:, :k].astype(np.float32), r64[:, :k], s64, r64)
    assert r["ok"] and r["overlap"] == 1.0, r
    del s32, rows


# ------------------------------------------------------------------------------------------
# seeded thresholds (shards of >= 2^20 rows): sample -> seed -> few big chunks -> verification
# ------------------------------------------------------------------------------------------

def _big(d=64, n=1_300_000, nq=96, seed=200):
    rng = np.random.Generator(np.random.PCG64(seed))
    xb = rng.standard_normal((n, d), dtype=np.float32)
    xq = rng.standard_normal((nq, d), dtype=np.float32)
    return xb, xq


@pytest.mark.parametrize("scan", ["f16", "tf32", "simt"])
def test_seeded_search_parity(cldrd_lib, scan):
    xb, xq = _big()
    gpu = _gpu_index(xb, None, scan)
    _check(gpu, xb, None, xq, 100)
    st = gpu.last_stats()
    # seeded pass: sample pieces + one chunk, far fewer survivors than the progressive scheme
    assert st["chunks"] <= 6, st
    assert st["survivors"] / xq.shape[0] < 40 * 100, st
    gpu.close()


def test_seed_miss_falls_back_and_stays_exact(cldrd_lib, monkeypatch):
    """A seed far above every score collects nothing: verification must flag every query and the
    unseeded retry must return the exact answer."""
    xb, xq = _big(nq=40, seed=201)
    monkeypatch.setenv("CLDRD_SEED_BIAS", "1e6")
    gpu = _gpu_index(xb, None, "f16")
    _check(gpu, xb, None, xq, 50)
    assert gpu.last_stats()["fallback_queries"] == xq.shape[0], gpu.last_stats()
    gpu.close()
    monkeypatch.setenv("CLDRD_SEED_BIAS", "0")
    monkeypatch.setenv("CLDRD_NO_SEED", "1")
    gpu = _gpu_index(xb, None, "f16")
    D0, I0 = gpu.search(xq, 50)
    assert gpu.last_stats()["fallback_queries"] == 0
    gpu.close()
    monkeypatch.setenv("CLDRD_NO_SEED", "0")
    gpu = _gpu_index(xb, None, "f16")
    D1, I1 = gpu.search(xq, 50)
    assert np.array_equal(D0, D1) and np.array_equal(I0, I1)   # seeded == progressive, to the bit
    gpu.close()


def test_sharded_seed_protocol_single_process(cldrd_lib):
    """The three-step protocol of cldrd.dist driven by hand over 3 shards on one GPU:
    sample -> seed from the gathered samples -> seeded search -> merge -> verify."""
    import torch
    from cldrd import dist as CD
    from cldrd._lib import check, lib, SEED_J
    from cldrd.index import shard_ranges
    xb, xq = _big(nq=64, seed=202)
    N = xb.shape[0]
    rows = torch.from_numpy(xb).cuda()
    q = torch.from_numpy(xq).cuda()
    k = 100
    shards = [CD.ShardedSearcher.from_rows(rows[rr.start:rr.stop], rr.start, N, scan="f16") for rr in shard_ranges(N, 3)]
    bound = max(_norm_bound(s) for s in shards)
    for s in shards:
        check(lib().cldrd_shard_set_norm_bound(s.shard.handle, C.c_float(bound)))
    topj = torch.stack([s.local.sample_device(q, k) for s in shards])
    assert topj.shape == (3, 64, SEED_J)
    seed = torch.empty((64,), dtype=torch.float32, device="cuda")
    check(lib().cldrd_seed_from_samples(0, C.c_void_p(topj.data_ptr()), 3, 64, C.c_void_p(seed.data_ptr()), None))
    outs = [s.local.search_device_seeded(q, k, seed) for s in shards]
    D, I = CD.merge_candidates(torch.stack([o[0] for o in outs]), torch.stack([o[1] for o in outs]))
    fail = torch.ones((64,), dtype=torch.int32, device="cuda")
    check(lib().cldrd_verify_seed(0, C.c_void_p(D.data_ptr()), 64, k, C.c_void_p(seed.data_ptr()),
                                  C.c_void_p(outs[0][2].data_ptr()), C.c_void_p(fail.data_ptr()), None))
    torch.cuda.synchronize()
    assert int(fail.sum()) <= 2          # the seed sits near rank 3k: misses are rare
    ok = (fail == 0).cpu().numpy()
    # the seed is a real filter: a shard collects ~J/f/3 rows per query, not its whole share
    assert shards[0].shard.stats()["survivors"] / 64 < 3000, shards[0].shard.stats()
    D_ref, I_ref = O.search(xb, None, xq, k)
    r = O.compare_topk(D.cpu().numpy()[ok], I.cpu().numpy()[ok], D_ref[ok], I_ref[ok],
                       *[a[ok] for a in O.search(xb, None, xq, k + 16, dtype=np.float64)])
    assert r["ok"], r
    # a hopeless seed is caught by the verification
    bad = seed + 1e6
    outs = [s.local.search_device_seeded(q, k, bad) for s in shards]
    D, I = CD.merge_candidates(torch.stack([o[0] for o in outs]), torch.stack([o[1] for o in outs]))
    check(lib().cldrd_verify_seed(0, C.c_void_p(D.data_ptr()), 64, k, C.c_void_p(bad.data_ptr()),
                                  C.c_void_p(outs[0][2].data_ptr()), C.c_void_p(fail.data_ptr()), None))
    torch.cuda.synchronize()
    assert int(fail.sum()) == 64


def _norm_bound(searcher):
    from cldrd._lib import check, lib
    b = C.c_float()
    check(lib().cldrd_shard_norm_bound(searcher.shard.handle, C.byref(b)))
    return b.value


def test_torchrun_two_ranks_nccl(cldrd_lib, tmp_path):
    """One process per GPU (needs 2 GPUs): sharded result == single-GPU result, bit for bit, through the
    peer-memory scatter (default) and through the NCCL all-to-all path."""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = tmp_path / "dist_result.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(root, "tests", "dist_worker.py"), "--out", str(out)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-8000:]
    import json
    res = json.loads(out.read_text())
    for name in ("small", "big", "manyq", "miss"):
        assert res[f"bit_equal_{name}_p2p"] and res[f"bit_equal_{name}_nccl"], res
        assert res[f"p2p_used_{name}"], res          # the peer-memory exchange is the path that ran
        assert res[f"bit_equal_{name}_host"] and res[f"host_shared_{name}"], res
    assert res["oracle_ok"] and res["host_owned_small"] and res["host_owned_big"] and res["on_batch_ok"], res
    assert res["seed_misses_miss"] == 33 and res["seed_misses_big"] == 0, res


def test_device_resident_queries_equal_numpy_path(cldrd_lib):
    """embeddings handed over on the device (no `.cpu().numpy()` round trip,
    retriever/retrieval_utils.py:47) give the same bits as the numpy call, on one GPU and on in-process shards."""
    import torch
    import cldrd
    from cldrd.retrieval_utils import index_retrieve_arrays
    xb, xq, ids = O.synth(40_000, 64, 420), O.synth(90, 64, 421), O.synth_ids(40_000, 422)
    gpu = _gpu_index(xb, ids, "f16")
    D, I = gpu.search(xq, 100)
    Dd, Id = index_retrieve_arrays(gpu, torch.from_numpy(xq).cuda(), 100)
    assert isinstance(Dd, np.ndarray) and np.array_equal(Dd, D) and np.array_equal(Id, I)
    with pytest.raises(TypeError):
        gpu.search(torch.from_numpy(xq).cuda().double(), 10)
    gpu.close()
    host = cldrd.IndexIDMap(cldrd.IndexFlatIP(64))
    host.add_with_ids(xb, ids)
    multi = cldrd.index_cpu_to_gpu_multiple(None, [0, 0], host, None)
    Dm, Im = multi.search(torch.from_numpy(xq).cuda(), 100)
    assert np.array_equal(Dm, D) and np.array_equal(Im, I)
    multi.close()


def test_more_queries_than_one_batch(cldrd_lib):
    """nq above the internal 8192-query batch (config 5 streams 502k queries the same way)."""
    xb, xq = O.synth(20_000, 64, 400), O.synth(8300, 64, 401)
    gpu = _gpu_index(xb, None, "f16")
    D, I = gpu.search(xq, 10)
    D_ref, I_ref = O.search(xb, None, xq, 10)
    r = O.compare_topk(D, I, D_ref, I_ref, *O.search(xb, None, xq, 26, dtype=np.float64))
    assert r["ok"], r
    gpu.close()


def test_bf16_scan_overlap_is_one(cldrd_lib):
    """BASELINE.json configs[3]: bf16 index scan + fp32 rescore vs the fp32-exact path: overlap@k = 1.0
    and identical bits, because both return the fp32 rescore of a provably complete candidate set."""
    xb, xq = O.synth(200_000, 768, 410), O.synth(64, 768, 411)
    res = {}
    for scan in ("bf16", "tf32"):
        gpu = _gpu_index(xb, None, scan)
        res[scan] = gpu.search(xq, 1000)
        gpu.close()
    assert np.array_equal(res["bf16"][1], res["tf32"][1]) and np.array_equal(res["bf16"][0], res["tf32"][0])


@pytest.mark.parametrize("nq", [4700, 8192])
def test_query_tile_counts_that_share_factors_with_the_grid(cldrd_lib, nq):
    """37 and 64 query tiles share factors with the 148-CTA grid: without the per-group rotation of
    the work units a CTA would only ever see some query tiles and their survivors would overflow a
    few (query, CTA) segments.  Checked against the SIMT scan (bit-equal) and: no fallback."""
    rng = np.random.Generator(np.random.PCG64(500))
    xb = rng.standard_normal((1_200_000, 64), dtype=np.float32)
    xq = rng.standard_normal((nq, 64), dtype=np.float32)
    gpu = _gpu_index(xb, None, "f16")
    D, I = gpu.search(xq, 100)
    st = gpu.last_stats()
    gpu.close()
    assert st["fallback_queries"] <= 2, st      # only genuine seed misses, if any
    ref = _gpu_index(xb, None, "simt")
    D0, I0 = ref.search(xq[:512], 100)
    ref.close()
    assert np.array_equal(D[:512], D0) and np.array_equal(I[:512], I0)


def test_hot_row_range_spills_to_the_pool_not_to_the_fallback(cldrd_lib):
    """3000 contiguous rows that every query likes (passages of one topic stored together): one CTA
    scans them in one or two work units and overflows its private survivor segment; the excess
    must land in the query's shared pool, not send the query to the fallback."""
    d = 64
    rng = np.random.Generator(np.random.PCG64(600))
    xb = rng.standard_normal((1_300_000, d), dtype=np.float32)
    topic = rng.standard_normal((d,), dtype=np.float32)
    topic /= np.linalg.norm(topic)
    xb[500_000:503_000] += 6.0 * topic[None, :]
    xq = (rng.standard_normal((64, d), dtype=np.float32) + 4.0 * topic[None, :]).astype(np.float32)
    gpu = _gpu_index(xb, None, "f16")
    _check(gpu, xb, None, xq, 1000)
    st = gpu.last_stats()
    assert st["fallback_queries"] == 0, st
    gpu.close()


@pytest.mark.parametrize("scan", ["tf32", "f16", "bf16"])
def test_error_bound_holds_on_heavy_tailed_values(cldrd_lib, scan):
    """The filter band is a worst-case bound c*|q|*|b|: check it on values with a huge dynamic range
    (log-normal magnitudes, random signs, exact cancellations), not just on N(0,1)."""
    import torch
    from cldrd._lib import check
    rng = np.random.Generator(np.random.PCG64(700))
    d = 1024

    def heavy(n):
        mag = np.exp(rng.normal(0.0, 2.5, size=(n, d))).astype(np.float32)
        return (mag * rng.choice([-1.0, 1.0], size=(n, d))).astype(np.float32)

    xb, xq = heavy(2048), heavy(256)
    xb[:64, 1::2] = -xb[:64, 0::2]          # pairs that cancel exactly
    xq[:32, 1::2] = xq[:32, 0::2]
    xb = np.clip(xb, -6e4, 6e4)             # stay inside fp16
    gpu = _gpu_index(xb, None, scan)
    q = torch.from_numpy(xq).cuda()
    out = torch.empty((256, 2048), dtype=torch.float32, device="cuda")
    check(cldrd_lib.cldrd_scan_dense_dev(gpu._shard.handle, C.c_void_p(q.data_ptr()), 256, 0, 2048, C.c_void_p(out.data_ptr()), None))
    torch.cuda.synchronize()
    got = out.cpu().numpy().astype(np.float64)
    ref = xq.astype(np.float64) @ xb.astype(np.float64).T
    scale = np.linalg.norm(xq.astype(np.float64), axis=1)[:, None] * np.linalg.norm(xb.astype(np.float64), axis=1)[None, :]
    rel = np.abs(got - ref) / scale
    coef = {"tf32": 2.0 ** -9, "f16": 2.0 ** -10, "bf16": 2.0 ** -7}[scan] + 2.2 * d * 2.0 ** -23
    assert rel.max() <= coef, (scan, rel.max(), coef)
    # and the search on such data is still exact
    D, I = gpu.search(xq[:40], 50)
    D_ref, I_ref = O.search(xb, None, xq[:40], 50)
    r = O.compare_topk(D, I, D_ref, I_ref, *O.search(xb, None, xq[:40], 66, dtype=np.float64), rel_tol=1e-4)
    assert r["bad_ids"] == 0, r
    gpu.close()


@pytest.mark.parametrize("scan", ["tf32", "f16", "bf16"])
@pytest.mark.parametrize("d", [768, 4096])
def test_accumulator_worst_case_same_sign_equal_magnitude(cldrd_lib, scan, d):
    """The one empirical term of the band: accum = 2.2*d*2^-23 (engine.cu eps_coefs) stands for whatever the tensor
    core does inside its accumulation.  Random signs hide a truncating accumulator (errors average out); this input
    is built to expose one: every operand is positive and exactly representable in bf16, fp16 AND tf32 (so operand
    rounding contributes nothing), the products are exact in fp32, but the partial sums need more than 24 bits, so
    every accumulation step has to round -- in the same direction if the adder truncates.  Measured against the
    exact value (integers in float64) and recorded under gpurun_out/ for profiles/."""
    import json
    import torch
    from cldrd._lib import check
    rng = np.random.Generator(np.random.PCG64(710 + d))
    nb, nq = 1024, 128
    # 8 significant bits each (bf16-exact); products near 4 with 14 fraction bits: already at d = 768 the partial sums
    # (up to ~3000) need 26 bits, at d = 4096 28 bits
    mags = np.array([2.0 - 2.0 ** -6, 2.0 - 2.0 ** -7, 1.5 + 2.0 ** -7, 2.0 - 2.0 ** -5], dtype=np.float32)
    xb = rng.choice(mags, size=(nb, d)).astype(np.float32)
    xq = rng.choice(mags, size=(nq, d)).astype(np.float32)
    xb[:256] = mags[0]                       # equal magnitudes everywhere: the purest same-direction case
    xq[:32] = mags[0]
    for a in (xb, xq):                       # the premise: nothing is lost when the operands are narrowed
        t = torch.from_numpy(a)
        assert torch.equal(t.bfloat16().float(), t) and torch.equal(t.half().float(), t)
    gpu = _gpu_index(xb, None, scan)
    assert gpu.scan == scan
    q = torch.from_numpy(xq).cuda()
    out = torch.empty((nq, nb), dtype=torch.float32, device="cuda")
    check(cldrd_lib.cldrd_scan_dense_dev(gpu._shard.handle, C.c_void_p(q.data_ptr()), nq, 0, nb, C.c_void_p(out.data_ptr()), None))
    torch.cuda.synchronize()
    got = out.cpu().numpy().astype(np.float64)
    ref = xq.astype(np.float64) @ xb.astype(np.float64).T          # exact: sums of multiples of 2^-14 below 2^53
    scale = np.linalg.norm(xq.astype(np.float64), axis=1)[:, None] * np.linalg.norm(xb.astype(np.float64), axis=1)[None, :]
    rel = np.abs(got - ref) / scale
    accum = 2.2 * d * 2.0 ** -23
    rec = {"scan": scan, "d": d, "max_err_over_qb_norms": float(rel.max()), "accum_term": accum,
           "ratio_to_accum_term": float(rel.max() / accum), "max_abs_err": float(np.abs(got - ref).max()),
           "signed_mean_err": float((got - ref).mean()), "equal_magnitude_block_max": float(rel[:32, :256].max()),
           "fp32_ulp_at_result": float(np.spacing(np.float32(ref.max())))}
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, f"accum_worst_case_{scan}_d{d}.json"), "w") as f:
            json.dump(rec, f)
    assert rel.max() <= accum, rec
    # and the whole search on such data (every score within a few ulps of its neighbours) is still exact
    D, I = gpu.search(xq[:16], 50)
    D_ref, I_ref = O.search(xb, None, xq[:16], 50)
    r = O.compare_topk(D, I, D_ref, I_ref, *O.search(xb, None, xq[:16], 66, dtype=np.float64))
    assert r["ok"], r
    gpu.close()


def test_sharded_searcher_from_file_and_in_process_shards_from_file(cldrd_lib, tmp_path):
    """File-backed paths: ShardedSearcher.from_file (each rank preads its own row range) and
    index_cpu_to_gpu_multiple on a lazily read index (two shards, rows streamed file -> HBM)."""
    import torch
    import cldrd
    from cldrd import dist as CD
    xb, xq, ids = O.synth(30_001, 96, 800), O.synth(77, 96, 801), O.synth_ids(30_001, 802) + 10 ** 10
    path = tmp_path / "ckpt.index"
    O.write_index(str(path), xb, ids)
    D_ref, I_ref = O.search(xb, ids, xq, 64)
    ext = O.search(xb, ids, xq, 80, dtype=np.float64)
    s = CD.ShardedSearcher.from_file(str(path), device=0, scan="f16")
    assert s.ntotal == 30_001 and s.id_map is not None
    D, I = s.search(torch.from_numpy(xq).cuda(), 64)
    r = O.compare_topk(D.cpu().numpy(), I.cpu().numpy(), D_ref, I_ref, *ext)
    assert r["ok"], r
    host = cldrd.read_index(str(path))
    assert host._rows.file is not None
    vres, vdev = cldrd.GpuResourcesVector(), cldrd.IntVector()
    for _ in range(2):
        vres.push_back(cldrd.StandardGpuResources())
        vdev.push_back(0)                      # two shards on the same device: exercises the merge path
    co = cldrd.GpuMultipleClonerOptions()
    co.shard = True
    multi = cldrd.index_cpu_to_gpu_multiple(vres, vdev, host, co)
    D2, I2 = multi.search(xq, 64)
    r = O.compare_topk(D2, I2, D_ref, I_ref, *ext)
    assert r["ok"], r
    assert np.array_equal(D2, D.cpu().numpy()) and np.array_equal(I2, I.cpu().numpy())
    multi.close()


@pytest.mark.parametrize("scan", ["f16", "tf32"])
def test_embedding_like_distribution_stays_exact_and_seeded(cldrd_lib, scan):
    """Encoder outputs are not white noise: a large mean vector shared by every passage and query and a
    decaying spectrum, so all scores crowd into a narrow range far from zero (|q||b| is ~10x the spread
    of the scores).  The filter band is wide relative to the score gaps here; results must still be exact,
    with no query sent to the fallback."""
    d, n, nq = 128, 1_200_000, 200
    rng = np.random.Generator(np.random.PCG64(900))
    mean = rng.standard_normal((d,), dtype=np.float32) * 0.8
    spectrum = (1.0 / np.sqrt(1.0 + np.arange(d, dtype=np.float32))).astype(np.float32)
    xb = (rng.standard_normal((n, d), dtype=np.float32) * spectrum[None, :] + mean[None, :]).astype(np.float32)
    xq = (rng.standard_normal((nq, d), dtype=np.float32) * spectrum[None, :] + mean[None, :]).astype(np.float32)
    gpu = _gpu_index(xb, None, scan)
    D, _ = _check(gpu, xb, None, xq, 1000)
    st = gpu.last_stats()
    assert st["fallback_queries"] == 0, st
    # the premise of the test: scores far from zero compared with their spread inside the top-k
    assert (D[:, 0] - D[:, -1]).max() < 0.2 * D[:, -1].min()
    gpu.close()


def _in_process_shards(rows, N, world, scan="f16", ids=None):
    """GpuIndexShards over `world` row ranges of a device tensor, all on GPU 0 (zero-copy shards)."""
    import cldrd
    from cldrd import dist as CD
    from cldrd.index import GpuIndexShards, shard_ranges
    parts = [CD.ShardedSearcher.from_rows(rows[rr.start:rr.stop], rr.start, N, scan=scan) for rr in shard_ranges(N, world)]
    return GpuIndexShards([p.shard for p in parts], ids, N, rows.shape[1])


@pytest.mark.parametrize("k", [100, 1000])
def test_node_protocol_three_shards_on_one_gpu(cldrd_lib, k):
    """The node-wide sharded search (cldrd_node_*) without processes: 3 shards on one GPU, each with its own
    exchange block and stream, blocks addressed directly.  sample -> levels -> scan -> counts -> counted cut ->
    re-score + scatter -> merge + seed check, all asynchronous, flag barriers between the shards' streams.
    Must equal the single-shard search bit for bit, and the counted cut must really cut."""
    import torch
    from cldrd import dist as CD
    xb, xq = _big(nq=70, seed=203)
    N, nq, world = xb.shape[0], xq.shape[0], 3
    ids = O.synth_ids(N, 204)
    rows = torch.from_numpy(xb).cuda()
    one = _gpu_index(xb, ids, "f16")
    D1, I1 = one.search(xq, k)
    one.close()
    multi = _in_process_shards(rows, N, world, ids=ids)
    D, I = multi.search(xq, k)
    assert multi.last_seed_misses <= 2
    assert np.array_equal(D, D1) and np.array_equal(I, I1)
    stats = multi.last_stats()
    assert all(st["tc_tiles"] > 0 for st in stats), stats
    rescored = sum(st["rescored"] for st in stats)
    # without the cut every shard re-scores all it collected above the seed (~3.5k rows per query over the
    # shards); with it, about k plus one level step plus the error band
    assert rescored / nq < 1.6 * k + 400, (rescored / nq, k)
    # results are caller-owned: a second search does not touch the first one's arrays
    D2, I2 = multi.search(xq[::-1].copy(), k)
    assert np.array_equal(D, D1) and np.array_equal(D2[::-1], D1) and np.array_equal(I2[::-1], I1)
    multi.close()


def test_node_protocol_retry_many_batches_and_small_shards(cldrd_lib, monkeypatch):
    """The rare paths of the node-wide search on one GPU: every seed missed (all queries raised by the merge's
    seed check and searched again unseeded), more queries than one 8192-query batch with several batches in
    flight and a query count that the shards do not divide, shards below the seeding size."""
    import torch
    xb, xq = _big(nq=33, seed=205)
    N = xb.shape[0]
    rows = torch.from_numpy(xb).cuda()
    one = _gpu_index(xb, None, "f16")
    D1, I1 = one.search(xq, 50)
    monkeypatch.setenv("CLDRD_SEED_BIAS", "1e6")
    multi = _in_process_shards(rows, N, 2)
    monkeypatch.setenv("CLDRD_SEED_BIAS", "0")
    D, I = multi.search(xq, 50)
    assert multi.last_seed_misses == 33
    assert np.array_equal(D, D1) and np.array_equal(I, I1)
    multi.close()
    # 3 full batches + a ragged one, k small; the ring of batches in flight wraps
    xq2 = O.synth(3 * 8192 + 77, 64, 206)
    multi = _in_process_shards(rows, N, 3)
    Dm, Im = multi.search(xq2, 10)
    Ds, Is = one.search(xq2, 10)
    assert np.array_equal(Dm, Ds) and np.array_equal(Im, Is)
    multi.close()
    one.close()
    # small index: unseeded batches (progressive scheme per shard), fewer rows than k on a shard
    xs, xqs = O.synth(2500, 64, 207), O.synth(40, 64, 208)
    rs = torch.from_numpy(xs).cuda()
    multi = _in_process_shards(rs, 2500, 3)
    one = _gpu_index(xs, None, "f16")
    for kk in (7, 1000, 2048):
        Dm, Im = multi.search(xqs, kk)
        Ds, Is = one.search(xqs, kk)
        assert np.array_equal(Dm, Ds) and np.array_equal(Im, Is), kk
    multi.close()
    one.close()


def test_shards_settle_on_one_scan_precision(cldrd_lib):
    """scan="auto" picks f16 or tf32 per shard from that shard's own value range; the error band of the sharded
    protocol must be the same on every shard, so one shard outside the fp16 range moves all of them to tf32."""
    import cldrd
    xb, xq = O.synth(40_000, 64, 210), O.synth(30, 64, 211)
    xb[35_000, 3] = 1.0e5                       # only the second shard exceeds the fp16 range
    host = cldrd.IndexFlatIP(64)
    host.add(xb)
    co = cldrd.GpuMultipleClonerOptions()
    co.shard = True
    co.scan = "auto"
    multi = cldrd.index_cpu_to_gpu_multiple(None, [0, 0], host, co)
    assert [sh.scan for sh in multi._shards] == ["tf32", "tf32"]
    D, I = multi.search(xq, 100)
    D_ref, I_ref = O.search(xb, None, xq, 100)
    r = O.compare_topk(D, I, D_ref, I_ref, *O.search(xb, None, xq, 116, dtype=np.float64))
    assert r["ok"], r
    multi.close()


def test_one_gpu_two_ranks_gloo_ipc(cldrd_lib, tmp_path):
    """cldrd.dist with one process per shard on a ONE-GPU box: two ranks on cuda:0, gloo for the setup
    collectives (NCCL refuses two ranks on one device), CUDA-IPC blocks on the same device.  The sharded result
    must equal the single-shard result bit for bit, through search and search_host."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = tmp_path / "dist_result.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(root, "tests", "dist_worker.py"), "--out", str(out), "--backend", "gloo",
           "--same-device"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-8000:]
    res = json.loads(out.read_text())
    for name in ("small", "big", "manyq", "miss"):
        assert res[f"bit_equal_{name}_p2p"], res
        assert res[f"p2p_used_{name}"], res          # the peer-memory exchange is the path that ran
        assert res[f"bit_equal_{name}_host"] and res[f"host_shared_{name}"], res
    assert res["oracle_ok"] and res["host_owned_small"] and res["host_owned_big"] and res["on_batch_ok"], res
    assert res["seed_misses_miss"] == 33 and res["seed_misses_big"] == 0, res
