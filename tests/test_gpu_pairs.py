"""GPU: svsb_top_pairs against the oracle's restatement of document_top_pairwise_scores' compute
(reference src/svs/kb.py:1650-1656 = np.dot(M, M.T) + get_top_pairs, src/svs/util.py:206-233).
fp32 scores: <= 1e-5 relative; pairs and ranks exact except across near-ties (oracle.compare_pairs)."""
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


def _unit(rng, shape, dist):
    m = rng.random(shape, dtype=np.float32) if dist == "uniform" else rng.standard_normal(shape).astype(np.float32)
    m /= np.maximum(np.sqrt((m * m).sum(axis=1)), 1e-12)[:, None]
    return m


@pytest.mark.parametrize("n_rows,d,n,dist", [
    (4, 3, 10, "normal"),               # the reference's own test scale: more pairs asked than exist (6)
    (300, 64, 50, "normal"),
    (1500, 96, 1000, "uniform"),
    (4875, 1536, 10000, "uniform"),     # the dad-jokes notebook's shape (Build Dad Jokes KB.ipynb:338-340)
    (5000, 256, 100, "normal"),         # bootstrap threshold + five query blocks
    (3001, 100, 1, "normal"),
])
def test_top_pairs_match_the_oracle(engine, n_rows, d, n, dist):
    rng = np.random.default_rng(n_rows + d)
    m = _unit(rng, (n_rows, d), dist)
    ids = np.cumsum(rng.integers(1, 4, size=n_rows)).astype(np.int64)
    engine.load(m, ids)
    got = engine.top_pairs(n)
    pairwise = np.dot(m, m.T)
    want = oracle.top_pairwise(m, ids, n)
    assert len(got) == min(n, n_rows * (n_rows - 1) // 2)
    rep = oracle.compare_pairs(got, want, pairwise, ids)
    assert rep["max_rel_score_err"] <= 1e-5


def test_top_pairs_with_duplicate_rows_orders_ties_by_rows(engine):
    rng = np.random.default_rng(3)
    m = _unit(rng, (2500, 48), "normal")
    m[10] = m[2000]; m[11] = m[2000]; m[700] = m[2000]            # four identical rows -> six pairs with score ~1
    ids = np.arange(100, 2600, dtype=np.int64)
    engine.load(m, ids)
    got = engine.top_pairs(8)
    assert [(a - 100, b - 100) for _, a, b in got[:6]] == [(10, 11), (10, 700), (10, 2000), (11, 700), (11, 2000), (700, 2000)]
    assert all(abs(s - 1.0) < 1e-5 for s, _, _ in got[:6])
    oracle.compare_pairs(got, oracle.top_pairwise(m, ids, 8), np.dot(m, m.T), ids)


def test_top_pairs_edge_cases(engine):
    rng = np.random.default_rng(4)
    m = _unit(rng, (50, 16), "normal")
    engine.load(m, np.arange(50, dtype=np.int64))
    assert engine.top_pairs(0) == []
    assert len(engine.top_pairs(10 ** 6)) == 50 * 49 // 2
    engine.load(m[:1], np.arange(1, dtype=np.int64))
    assert engine.top_pairs(5) == []
