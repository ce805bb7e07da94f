"""GPU: small exact batches -- gemv_tma_mq_kernel streams the matrix ONCE for up to 8 warps' worth of queries, then
one selection CTA per query (svsb_query_batch for 2 <= b <= SVSB_MQ_MAX, and whenever the tensor-core path is not
available: tombstones, k > 1024, tiny matrices).  The promise is the SAME BITS as b calls of svsb_query: every
(row, query) dot product is computed with the single-query kernel's summation order."""
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


def _same_as_single(engine, q, k, s, i, c):
    n = engine.shape[0]
    assert (c == min(k, n)).all()
    for j in range(len(q)):
        ss, ii = engine.query(q[j], k)
        assert np.array_equal(ss.view(np.uint32), s[j, :len(ss)].view(np.uint32)), f"scores differ for query {j}"
        assert np.array_equal(ii, i[j, :len(ii)]), f"ids differ for query {j}"


@pytest.mark.parametrize("d", [64, 100, 257, 768, 1536, 1537, 3072])
def test_every_batch_size_equals_single_query_bits(engine, d, monkeypatch):
    import svs_b200
    monkeypatch.setenv("SVSB_MQ_MAX", "64")                        # also where the tensor-core path would be preferred
    n = 30_011 if d <= 1536 else 9_001                            # ragged last tile
    rng = np.random.default_rng(d)
    m = rng.standard_normal((n, d)).astype(np.float32)
    m /= np.sqrt((m * m).sum(axis=1))[:, None]
    ids = np.cumsum(rng.integers(1, 4, size=n)).astype(np.int64)
    engine.load(m, ids)
    qs = oracle.synth_queries(40, d, 3, dist="normal")
    for b in (2, 3, 4, 5, 7, 8, 9, 15, 16, 17, 31, 32):
        k = (1, 10, 100, 2048)[b % 4]
        l0 = svs_b200.launch_count()
        s, i, c = engine.query_batch(qs[:b], k)
        launches = svs_b200.launch_count() - l0
        _same_as_single(engine, qs[:b], k, s, i, c)
        if b <= 4:
            assert launches == 2, f"b={b}: one multi-query similarity pass + one batched selection expected, saw {launches} launches"
    # and against the oracle, once
    s, i, c = engine.query_batch(qs[:6], 50)
    for j in range(6):
        got = list(zip(s[j].tolist(), i[j].tolist()))
        oracle.compare_retrieval(got, oracle.superheavy(m, ids, qs[j], 50), oracle.scores_of(m, qs[j]), ids)


def test_rows_too_long_for_register_resident_queries_fall_back(engine):
    d = 3200                                                       # > 3072: the exact path loops single queries (or the coarse pass)
    m = oracle.synth_matrix_normal(5000, d, 1)
    engine.load(m, np.arange(5000, dtype=np.int64))
    qs = oracle.synth_queries(5, d, 2, dist="normal")
    s, i, c = engine.query_batch(qs, 10)
    _same_as_single(engine, qs, 10, s, i, c)


def test_tiny_matrices_ties_and_edge_ks(engine):
    qs = np.ones((6, 4), dtype=np.float32)
    engine.load(np.ones((1000, 4), dtype=np.float32), np.arange(1000, dtype=np.int64))
    s, i, c = engine.query_batch(qs, 7)
    assert all(i[j].tolist() == list(range(7)) for j in range(6))  # all scores tie: ascending id
    for rows in (1, 2, 5):
        m = oracle.synth_matrix_normal(rows, 16, rows)
        engine.load(m, np.arange(10, 10 + rows, dtype=np.int64))
        q = oracle.synth_queries(3, 16, 4, dist="normal")
        s, i, c = engine.query_batch(q, 4)
        _same_as_single(engine, q, 4, s, i, c)
        assert engine.query_batch(q, 0)[2].tolist() == [0, 0, 0]
    with pytest.raises(ValueError):
        engine.query_batch(np.zeros((3, 17), np.float32), 3)


def test_generations_with_tombstones_answer_batches_exactly(engine):
    """The tensor-core path is off for a generation with tombstones; batches of any size then run as multi-query passes."""
    n, d = 20_000, 384
    m = oracle.synth_matrix_normal(n, d, 5)
    ids = np.arange(1, n + 1, dtype=np.int64)
    engine.load(m, ids)
    qs = oracle.synth_queries(70, d, 6, dist="normal")
    before = engine.query_batch(qs[:5], 20)
    dead = np.unique(np.concatenate([before[1][:, :3].reshape(-1), np.arange(100, 400)]))
    engine.apply_mutations(dead, [], None)
    keep = ~np.isin(ids, dead)
    for b in (5, 70):
        s, i, c = engine.query_batch(qs[:b], 20)
        _same_as_single(engine, qs[:b], 20, s, i, c)
        assert not np.isin(i, dead).any()
    got = list(zip(s[0].tolist(), i[0].tolist()))
    oracle.compare_retrieval(got, oracle.superheavy(m[keep], ids[keep], qs[0], 20), oracle.scores_of(m[keep], qs[0]), ids[keep])
