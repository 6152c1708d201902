"""The exactness argument of the sharded search (DESIGN.md sections 4, 5, 7) attacked on the CPU: scan scores
that differ from the exact scores by the full error bound in the most damaging directions must never cost a
verified query one of its true top-k rows, with or without the level-count cut, whatever counts get lost."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
from oracle import sharded_filter as SF  # noqa: E402


def _scores(nq, N, seed, spread=1.0, offset=0.0):
    rng = np.random.Generator(np.random.PCG64(seed))
    return (rng.standard_normal((nq, N)) * spread + offset).astype(np.float32)


def _errors(S, eps, kind, k, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    if kind == "random":
        return rng.uniform(-eps, eps, size=S.shape).astype(np.float32)
    if kind == "sign":
        return (rng.integers(0, 2, size=S.shape) * 2 - 1).astype(np.float32) * np.float32(eps)
    # adversarial: push every true top-k row DOWN by eps and everything else UP by eps
    E = np.full(S.shape, eps, dtype=np.float32)
    top = np.argsort(-S, axis=1)[:, :k]
    np.put_along_axis(E, top, -eps, axis=1)
    return E


@pytest.mark.parametrize("shards", [1, 2, 3, 8])
@pytest.mark.parametrize("kind", ["random", "sign", "adversarial"])
@pytest.mark.parametrize("k", [10, 100])
def test_verified_queries_keep_their_true_top_k(shards, kind, k):
    S = _scores(12, 60_000, 10 * shards + k)
    eps = 0.02
    S_hat = (S + _errors(S, eps, kind, k, 7)).astype(np.float32)
    eps_eff = float(np.abs(S_hat.astype(np.float64) - S).max())        # float32 rounding of the sum
    res = SF.run(S, S_hat, eps_eff, k, shards, sample_stride=16)
    assert res["verified"].sum() >= 10                                  # the seed sits far below rank k
    assert res["ok"][res["verified"]].all(), (shards, kind, k, res)
    base = SF.run(S, S_hat, eps_eff, k, shards, sample_stride=16, use_cut=False)
    assert base["ok"][base["verified"]].all()
    assert (res["rescored"] <= base["rescored"]).all()


def test_cut_shrinks_the_rescored_set_to_about_k():
    k, shards = 100, 8
    S = _scores(16, 200_000, 3)
    eps = 0.005
    S_hat = (S + _errors(S, eps, "random", k, 4)).astype(np.float32)
    eps_eff = float(np.abs(S_hat.astype(np.float64) - S).max())
    with_cut = SF.run(S, S_hat, eps_eff, k, shards, sample_stride=64)
    without = SF.run(S, S_hat, eps_eff, k, shards, sample_stride=64, use_cut=False)
    assert with_cut["ok"][with_cut["verified"]].all()
    # the seed sits near rank J * stride = 2048: that many rows are collected over the shards.  Without the
    # cut every shard re-scores its own best k (plus band) of them: ~8 * k in total.  With it: k plus one
    # level step (stride rows) plus the band.
    assert with_cut["collected"].mean() > 15 * k
    assert without["rescored"].mean() > 7 * k
    assert with_cut["rescored"].mean() < 2.5 * k


def test_lost_counts_only_cost_work_never_rows():
    k, shards = 50, 4
    S = _scores(20, 80_000, 5)
    eps = 0.01
    S_hat = (S + _errors(S, eps, "adversarial", k, 6)).astype(np.float32)
    eps_eff = float(np.abs(S_hat.astype(np.float64) - S).max())
    rng = np.random.Generator(np.random.PCG64(9))
    res = SF.run(S, S_hat, eps_eff, k, shards, sample_stride=32, lost_counts=0.5, rng=rng)
    full = SF.run(S, S_hat, eps_eff, k, shards, sample_stride=32)
    assert res["ok"][res["verified"]].all()
    assert (res["rescored"] >= full["rescored"]).all()                  # fewer counts -> lower cut -> more work


def test_narrow_score_range_far_from_zero():
    """Embedding-like scores: everything near 95 +- 2 with an error band that is wide relative to the gaps."""
    k, shards = 100, 2
    S = _scores(8, 100_000, 11, spread=2.0, offset=95.0)
    eps = 0.12
    S_hat = (S + _errors(S, eps, "sign", k, 12)).astype(np.float32)
    eps_eff = float(np.abs(S_hat.astype(np.float64) - S).max())
    res = SF.run(S, S_hat, eps_eff, k, shards, sample_stride=32)
    assert res["ok"][res["verified"]].all() and res["verified"].sum() >= 6


def test_seed_above_the_true_kth_is_caught_by_verification():
    k = 20
    S = _scores(6, 20_000, 13)
    S_hat = S.copy()
    # stride so large that fewer than J samples exist per shard -> seed = -inf (unseeded), always verified
    res = SF.run(S, S_hat, 1e-6, k, 2, sample_stride=5000)
    assert res["verified"].all() and res["ok"].all()
    # J = 4 levels from a tiny sample: the seed lands ABOVE the k-th score for most queries; those must fail
    res = SF.run(S, S_hat, 1e-6, k, 2, J=4, sample_stride=2)
    assert (~res["verified"]).sum() >= 4
    assert res["ok"][res["verified"]].all()


def test_levels_are_the_top_j_of_the_union():
    rng = np.random.Generator(np.random.PCG64(1))
    topj = -np.sort(-rng.standard_normal((3, 5, 8)).astype(np.float32), axis=2)
    topj[2, :, 5:] = -np.inf
    lv = SF.levels_from_samples(topj)
    for q in range(5):
        exp = np.sort(np.concatenate([topj[p, q] for p in range(3)]))[::-1][:8]
        assert np.array_equal(lv[q], exp)


def test_merge_by_rank_equals_sort_of_the_union():
    """The merge kernel of the node-wide search (csrc/node.cuh: merge_keys_kernel) does not sort: every shard's list
    arrives sorted best-first, and a key's place in the merged order is its place in its own list plus, for every other
    list, the number of keys there that order before it (one binary search per list; equal keys of different shards go
    lower shard first).  Restated in numpy and compared with a plain sort of the union, on random lists of ragged
    lengths, empty lists and forced cross-shard ties."""
    rng = np.random.Generator(np.random.PCG64(77))
    for trial in range(200):
        parts = int(rng.integers(1, 9))
        k = int(rng.integers(1, 40))
        lists = []
        for p in range(parts):
            n = int(rng.integers(0, k + 1))
            keys = rng.integers(1, 60, size=n).astype(np.uint64)            # small range: ties across shards happen
            keys = np.unique(keys)[::-1][:n]                                # a shard's own keys are distinct, descending
            lists.append(keys)
        merged = np.zeros(k, dtype=np.uint64)
        filled = np.zeros(k, dtype=bool)
        for p, mine in enumerate(lists):
            for j, key in enumerate(mine):
                rank = j
                for o, other in enumerate(lists):
                    if o == p:
                        continue
                    before = other > key if o > p else other >= key         # ties: the lower shard goes first
                    rank += int(before.sum())
                if rank < k:
                    assert not filled[rank], "two keys claimed one place"
                    merged[rank], filled[rank] = key, True
        union = np.concatenate([np.stack([l, np.full(l.shape, p, dtype=np.uint64)], 1) for p, l in enumerate(lists)] or
                               [np.zeros((0, 2), dtype=np.uint64)])
        order = np.lexsort((union[:, 1], -union[:, 0].astype(np.int64)))    # key descending, shard ascending
        want = union[order][:k, 0]
        m = want.shape[0]
        assert filled[:m].all() and not filled[m:].any()
        assert np.array_equal(merged[:m], want)
