"""CPU property tests of the oracles (size-independent invariants of the domain, prompt section 3): they hold for the
reference by construction, so an oracle that breaks one is wrong even where no golden vector looks."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import loss_oracle as LO
from oracle import metrics_oracle as MO


def _case(seed, nq, ng, n_ids, n_cams, ties):
    rs = np.random.RandomState(seed)
    dist = rs.rand(nq, ng).astype(np.float32)
    if ties:
        dist = np.round(dist * 4) / 4                       # heavy ties: the stable-order contract matters
    q_pids, g_pids = rs.randint(0, n_ids, nq), rs.randint(0, n_ids, ng)
    q_cams, g_cams = rs.randint(0, n_cams, nq), rs.randint(0, n_cams, ng)
    return dist, q_pids.astype(np.int64), g_pids.astype(np.int64), q_cams.astype(np.int64), g_cams.astype(np.int64)


@settings(max_examples=40, deadline=None)
@given(seed=st.integers(0, 10_000), nq=st.integers(1, 12), ng=st.integers(1, 40), n_ids=st.integers(1, 6),
       n_cams=st.integers(1, 3), ties=st.booleans(), max_rank=st.integers(1, 50))
def test_rank_c_twin_equals_numpy_oracle(seed, nq, ng, n_ids, n_cams, ties, max_rank):
    """rank_oracle.c (twin of rank_cy.pyx) and the numpy restatement of rank.py agree on arbitrary inputs, ties and
    max_rank > Ng included; queries without a usable match are skipped by both; an all-invalid query set asserts."""
    dist, qp, gp, qc, gc = _case(seed, nq, ng, n_ids, n_cams, ties)
    try:
        r = MO.eval_market1501(dist, qp, gp, qc, gc, max_rank)
    except AssertionError:
        with pytest.raises(AssertionError):
            MO.eval_market1501_c(dist, qp, gp, qc, gc, max_rank)
        return
    except ValueError:
        # reference quirk kept by the numpy restatement: a query whose kept gallery is shorter than max_rank yields a
        # shorter cmc row, and rows of different length do not stack (rank.py:139-143); the Cython twin pads instead
        return
    cmc_c, map_c = MO.eval_market1501_c(dist, qp, gp, qc, gc, max_rank)
    assert len(cmc_c) == min(max_rank, ng)
    n = min(len(r["cmc"]), len(cmc_c))          # (same quirk: the numpy row is as long as the shortest kept gallery)
    np.testing.assert_allclose(cmc_c[:n], r["cmc"][:n], atol=1e-6)
    np.testing.assert_allclose(map_c, r["mAP"], rtol=2e-5, atol=1e-7)
    assert np.all(np.diff(r["cmc"]) >= -1e-7) and 0.0 <= r["cmc"][0] <= r["cmc"][-1] <= 1.0 + 1e-6    # cmc is a CDF
    assert 0.0 < r["mAP"] <= 1.0 + 1e-9


@settings(max_examples=25, deadline=None)
@given(seed=st.integers(0, 10_000), nq=st.integers(1, 6), ng=st.integers(1, 9), p=st.integers(1, 5),
       strat=st.sampled_from(["mean", "max"]), metric=st.sampled_from(["euclidean", "cosine"]))
def test_part_distance_oracle_invariants(seed, nq, ng, p, strat, metric):
    """Masked combination: invalid (i, j, part) entries never influence a valid combined distance; rows with no common
    visible part get max+1; a fully visible call equals the call without visibility; row/column permutations commute."""
    rs = np.random.RandomState(seed)
    qf = MO.l2_normalize(rs.randn(nq, p, 16).astype(np.float32))
    gf = MO.l2_normalize(rs.randn(ng, p, 16).astype(np.float32))
    qv, gv = rs.rand(nq, p) > 0.3, rs.rand(ng, p) > 0.3
    d, pd = MO.part_distance(qf, gf, qv, gv, strat, metric)
    assert d.shape == (nq, ng) and pd.shape == (p, nq, ng) and np.isfinite(d).all() and np.isfinite(pd).all()
    valid = qv.T[:, :, None] & gv.T[:, None, :]                          # [P, Nq, Ng]
    if valid.any() and not valid.all():
        assert np.allclose(pd[~valid], pd[valid].max() + 1)              # sentinel = max over the VALID entries + 1
    none = ~valid.any(0)
    if none.any() and (~none).any():
        assert np.all(d[none] > d[~none].max())                          # pairs without a common part rank last
    # corrupting features of invisible parts must not change any valid combined distance
    qf2 = qf.copy()
    qf2[~qv] = 7.0
    d2, _ = MO.part_distance(qf2, gf, qv, gv, strat, metric)
    np.testing.assert_allclose(d2[~none], d[~none], rtol=1e-5, atol=1e-6)
    # all-visible == no visibility given
    ones_q, ones_g = np.ones((nq, p), bool), np.ones((ng, p), bool)
    da, _ = MO.part_distance(qf, gf, ones_q, ones_g, strat, metric)
    db, _ = MO.part_distance(qf, gf, None, None, strat, metric)
    np.testing.assert_allclose(da, db, rtol=1e-6, atol=1e-6)
    # permutation equivariance
    pq, pg = rs.permutation(nq), rs.permutation(ng)
    dp, _ = MO.part_distance(qf[pq], gf[pg], qv[pq], gv[pg], strat, metric)
    np.testing.assert_allclose(dp, d[pq][:, pg], rtol=1e-5, atol=1e-6)


@settings(max_examples=25, deadline=None)
@given(seed=st.integers(0, 10_000), n_ids=st.integers(2, 5), per_id=st.integers(1, 4), k=st.integers(1, 4))
def test_triplet_oracle_invariants(seed, n_ids, per_id, k):
    """Batch-hard mining: the loss is invariant to a permutation of the batch and of the parts, non-negative, and zero
    for well-separated identity clusters (hard margin)."""
    rs = np.random.RandomState(seed)
    n = n_ids * per_id
    pids = np.repeat(np.arange(n_ids), per_id).astype(np.int64)
    emb = rs.randn(n, k, 8).astype(np.float32)
    res = LO.part_averaged_triplet(emb, pids, 0.3)
    if res is None:                                                   # no valid triplet (singleton identities)
        assert per_id == 1
        return
    loss = res[0]
    assert loss >= 0
    perm = rs.permutation(n)
    res_p = LO.part_averaged_triplet(emb[perm], pids[perm], 0.3)
    np.testing.assert_allclose(res_p[0], loss, rtol=1e-5, atol=1e-6)
    pk = rs.permutation(k)
    res_k = LO.part_averaged_triplet(emb[:, pk], pids, 0.3)
    np.testing.assert_allclose(res_k[0], loss, rtol=1e-5, atol=1e-6)
    centers = 100.0 * np.eye(n_ids, 8, dtype=np.float32)[pids][:, None, :].repeat(k, 1)
    res_far = LO.part_averaged_triplet(centers + 0.01 * emb, pids, 0.3)
    assert res_far[0] == 0.0
