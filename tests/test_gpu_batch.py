"""GPU: svsb_query_batch (tensor-core coarse contraction + exact fp32 refine) against the single-query path and the
oracle.  The batched path promises the SAME BITS as b calls of svsb_query, so scores are
compared as uint32 and ids exactly; one configuration is also checked against the oracle's superheavy()
(reference src/svs/kb.py:1622-1627) under the usual tolerance."""
import numpy as np
import pytest

from _util import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    import svs_b200
    e = svs_b200.Engine()
    yield e
    e.close()


@pytest.fixture(autouse=True)
def _coarse_path_for_every_batch_size(monkeypatch):
    """This file tests the tensor-core path; small batches would otherwise take the exact multi-query passes
    (tests/test_gpu_mq.py), which answer up to SVSB_MQ_MAX queries when both paths are available."""
    monkeypatch.setenv("SVSB_MQ_MAX", "1")


def _unit(rng, shape, dist):
    m = rng.random(shape, dtype=np.float32) if dist == "uniform" else rng.standard_normal(shape).astype(np.float32)
    m /= np.maximum(np.sqrt((m * m).sum(axis=1)), 1e-12)[:, None]
    return m


def _assert_same_as_single(engine, q, k, s, i, c, check=None):
    n = engine.shape[0]
    assert (c == min(k, n)).all()
    for j in (range(len(q)) if check is None else check):
        ss, ii = engine.query(q[j], k)
        assert np.array_equal(ss.view(np.uint32), s[j, :len(ss)].view(np.uint32)), f"scores differ for query {j}"
        assert np.array_equal(ii, i[j, :len(ii)]), f"ids differ for query {j}"


@pytest.mark.parametrize("n,d,k,b,dist", [
    (20_000, 768, 100, 300, "uniform"),        # the README recipe's distribution: scores 0.75 +- 0.007, a hard case
    (20_000, 768, 100, 300, "normal"),
    (5_000, 96, 10, 8, "normal"),              # one partial query tile
    (70_001, 1536, 100, 64, "uniform"),        # ragged last row tile
    (30_000, 100, 7, 33, "normal"),            # d not a multiple of 8
    (12_345, 3072, 1000, 20, "uniform"),       # large k
    (40_000, 256, 1, 513, "normal"),           # k = 1, three query tiles
    (300_000, 128, 10, 64, "normal"),          # sample = 1.4 % of the rows: statistical threshold (7th largest of the sample)
    (400_000, 64, 100, 40, "uniform"),         # sample = 8 % of the rows: statistical threshold (29th largest)
])
def test_batch_equals_single_query_bits(engine, n, d, k, b, dist):
    rng = np.random.default_rng(n + d + k)
    m = _unit(rng, (n, d), dist)
    ids = np.cumsum(rng.integers(1, 4, size=n)).astype(np.int64)
    q = _unit(rng, (b, d), dist)
    engine.load(m, ids)
    s, i, c = engine.query_batch(q, k)
    cand, resc, flags = engine.batch_stats(b)
    assert (flags == 0).all(), "these inputs must be answered by the coarse path, not the fallback"
    assert (resc >= min(k, n)).all() and (cand >= resc).all()
    _assert_same_as_single(engine, q, k, s, i, c, check=range(0, b, max(1, b // 24)))


def test_batch_against_the_oracle(engine):
    rng = np.random.default_rng(5)
    n, d, k, b = 50_000, 768, 100, 16
    m = _unit(rng, (n, d), "uniform")
    ids = np.arange(1, n + 1, dtype=np.int64)
    q = _unit(rng, (b, d), "uniform")
    engine.load(m, ids)
    s, i, c = engine.query_batch(q, k)
    for j in range(b):
        got = [(float(a), int(x)) for a, x in zip(s[j, :c[j]], i[j, :c[j]])]
        rep = oracle.compare_retrieval(got, oracle.superheavy(m, ids, q[j], k), oracle.scores_of(m, q[j]), ids)
        assert rep["max_rel_score_err"] <= 1e-5


def test_batch_with_massive_ties_falls_back_and_stays_exact(engine):
    """Every row identical: all coarse scores tie, candidate lists overflow, the exact kernels answer."""
    rng = np.random.default_rng(9)
    n, d, k, b = 40_000, 128, 50, 8
    row = _unit(rng, (1, d), "normal")
    m = np.repeat(row, n, axis=0)
    ids = np.arange(n, dtype=np.int64)
    q = _unit(rng, (b, d), "normal")
    engine.load(m, ids)
    s, i, c = engine.query_batch(q, k)
    cand, resc, flags = engine.batch_stats(b)
    assert (flags != 0).all()
    assert (i == np.arange(k)[None, :]).all()                     # ties: ascending row / id
    _assert_same_as_single(engine, q, k, s, i, c)


def test_batch_with_thousands_of_survivors_spills_and_stays_exact(engine):
    """2 500 rows within a hair of each other at the top of every query's ranking: all of them lie inside the 2-eps band
    under the k-th coarse score, so the refine kernel must re-score ~2 500 rows per query -- more than it holds in shared
    memory (1024): the rest goes through its global lists.  No query may be handed to the fallback, and the bits must
    equal the single-query path's."""
    rng = np.random.default_rng(321)
    n, d, k, b = 60_000, 128, 100, 24
    m = _unit(rng, (n, d), "normal")
    centre = _unit(rng, (1, d), "normal")[0]
    near = centre[None, :] + 2e-4 * rng.standard_normal((2500, d)).astype(np.float32)
    near /= np.sqrt((near * near).sum(axis=1))[:, None]
    where = rng.choice(n, size=2500, replace=False)
    m[where] = near
    ids = np.arange(10, n + 10, dtype=np.int64)
    q = centre[None, :] + 0.05 * rng.standard_normal((b, d)).astype(np.float32)
    q /= np.sqrt((q * q).sum(axis=1))[:, None]
    q = q.astype(np.float32)
    engine.load(m, ids)
    s, i, c = engine.query_batch(q, k)
    cand, resc, flags = engine.batch_stats(b)
    assert (flags == 0).all(), flags
    assert resc.min() > 1024 and resc.max() <= 4096, (resc.min(), resc.max())
    _assert_same_as_single(engine, q, k, s, i, c)


def test_statistical_thresholds_fail_verification_on_sorted_rows_and_stay_exact(engine):
    """Rows stored in descending similarity to the queries' common direction: the strided row sample then holds the
    overall top rows, the sample's order statistic lands ABOVE the true cut-off, the refine kernel's verification
    (flag 16 / 4) catches it, the chunk is redone with proven-bound thresholds and the generation stays in that mode.
    Results must still be the single-query path's bits."""
    rng = np.random.default_rng(21)
    n, d, k, b = 300_000, 64, 20, 64
    base = _unit(rng, (1, d), "normal")[0]
    m = _unit(rng, (n, d), "normal")
    m = m[np.argsort(-(m @ base), kind="stable")]                 # best rows first: row tile 0 is always sampled
    ids = np.arange(1, n + 1, dtype=np.int64)
    q = base[None, :] + 0.05 * rng.standard_normal((b, d)).astype(np.float32)
    q /= np.sqrt((q * q).sum(axis=1))[:, None]
    engine.load(m, ids)
    assert engine.batch_threshold_mode() == 0
    s, i, c = engine.query_batch(q, k)
    assert engine.batch_threshold_mode() == 1, "verification should have failed for most queries"
    cand, resc, flags = engine.batch_stats(b)
    assert (flags == 0).all()                                      # answered by the redone coarse pass, not one by one
    _assert_same_as_single(engine, q, k, s, i, c, check=range(0, b, 4))
    s2, i2, c2 = engine.query_batch(q, k)                          # later batches go straight to the proven bound
    assert np.array_equal(s.view(np.uint32), s2.view(np.uint32)) and np.array_equal(i, i2)
    engine.load(m[::-1].copy(), ids)                               # a new generation starts statistical again
    assert engine.batch_threshold_mode() == 0


def test_batch_edge_cases(engine):
    rng = np.random.default_rng(11)
    n, d = 6_000, 64
    m = _unit(rng, (n, d), "normal")
    engine.load(m, np.arange(n, dtype=np.int64))
    q = _unit(rng, (10, d), "normal")
    s, i, c = engine.query_batch(q, 0)                             # k <= 0 -> no results (util.py:200-201)
    assert (c == 0).all() and s.shape == (10, 0)
    s, i, c = engine.query_batch(q[:2], 5)                         # tiny batch: loops the single-query kernels
    _assert_same_as_single(engine, q[:2], 5, s, i, c)
    s, i, c = engine.query_batch(q, 3000)                          # k above the coarse path's limit
    _assert_same_as_single(engine, q, 3000, s, i, c, check=[0, 9])
    with pytest.raises(ValueError):
        engine.query_batch(_unit(rng, (10, d + 1), "normal"), 5)   # d mismatch, as np.dot
    big = q.copy(); big[3] *= 100.0                                # a query far from unit norm: exact path for it only
    s, i, c = engine.query_batch(big, 5)
    cand, resc, flags = engine.batch_stats(10)
    assert flags[3] != 0 and (np.delete(flags, 3) == 0).all()
    _assert_same_as_single(engine, big, 5, s, i, c)


def test_batch_with_page_locked_buffers_matches_the_staged_path(engine):
    """Page-locked query / result arrays (svsb_host_alloc) are DMA'd directly; pageable ones are staged: same bits.
    Covers two chunks (b > 2048) and a flagged query whose row is rewritten by the exact path in place."""
    from svs_b200 import pinned_empty
    rng = np.random.default_rng(17)
    n, d, k, b = 30_000, 256, 20, 2100
    m = _unit(rng, (n, d), "normal")
    engine.load(m, np.arange(5, 5 + n, dtype=np.int64))
    q = _unit(rng, (b, d), "normal")
    q[7] *= 50.0                                                   # far from unit norm: exact path for this query
    hq = pinned_empty((b, d), np.float32)
    hq[:] = q
    out = (pinned_empty((b, k), np.float32), pinned_empty((b, k), np.int64), np.zeros(b, dtype=np.int32))
    s1, i1, c1 = engine.query_batch(hq, k, out=out)
    assert s1 is out[0] and i1 is out[1]
    s2, i2, c2 = engine.query_batch(q, k)
    assert np.array_equal(s1.view(np.uint32), s2.view(np.uint32)) and np.array_equal(i1, i2) and np.array_equal(c1, c2)
    _assert_same_as_single(engine, q, k, s1, i1, c1, check=[0, 7, 2047, 2048, 2099])
    with pytest.raises(ValueError):
        engine.query_batch(q, k, out=(out[0][:10], out[1], out[2]))


def test_batch_after_reload_uses_the_new_generation(engine):
    rng = np.random.default_rng(13)
    d, k, b = 128, 10, 16
    q = _unit(rng, (b, d), "normal")
    for n in (8_000, 9_000):
        m = _unit(rng, (n, d), "normal")
        engine.load(m, np.arange(100, 100 + n, dtype=np.int64))
        s, i, c = engine.query_batch(q, k)
        _assert_same_as_single(engine, q, k, s, i, c)
