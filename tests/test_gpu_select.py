"""GPU: the selection kernels (K3/K4) against get_top_k -- reference src/svs/util.py:190-203 -- called
through the C ABI (svsb_topk_scores).  Integer/index work: bit-exact."""
import numpy as np
import pytest

from _util import golden_json, golden_npz, oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    import svs_b200
    e = svs_b200.Engine()
    yield e
    e.close()


def _check_exact(engine, scores, k):
    """Engine result == exact top-k under (score desc, index asc), scores bit-identical."""
    s32 = np.ascontiguousarray(scores, dtype=np.float32)
    got = engine.topk_scores(s32, k)
    want = oracle.canonical_top_k(s32, np.arange(len(s32), dtype=np.int64), k)
    assert len(got) == len(want)
    assert [i for _, i in got] == [i for _, i in want]
    assert np.array([s for s, _ in got], dtype=np.float32).tobytes() == np.array([s for s, _ in want], dtype=np.float32).tobytes()
    return got


def test_reference_known_answers(engine):
    # tests/test_util.py:142-400 of the reference (distinct scores: order is unambiguous)
    assert engine.topk_scores(np.array([]), 0) == []
    assert engine.topk_scores(np.array([]), 1) == []
    assert engine.topk_scores(np.array([0.4]), 0) == []
    f = lambda v: float(np.float32(v))
    assert engine.topk_scores(np.array([0.4]), 1) == [(f(0.4), 0)]
    assert engine.topk_scores(np.array([0.4]), 2) == [(f(0.4), 0)]
    assert engine.topk_scores(np.array([0.4, 0.2]), 2) == [(f(0.4), 0), (f(0.2), 1)]
    assert engine.topk_scores(np.array([0.2, 0.4]), 1) == [(f(0.4), 1)]
    assert engine.topk_scores(np.array([0.2, 0.4]), 3) == [(f(0.4), 1), (f(0.2), 0)]
    assert engine.topk_scores(np.array([0.2, 0.4]), -5) == []


def test_golden_cases_match_the_reference(engine):
    arrays = golden_npz("topk_cases.npz")
    n = 0
    for c in golden_json("topk_cases.json"):
        a = arrays[c["scores"]]
        a32 = a.astype(np.float32)
        if len(np.unique(a32)) != len(a32) or (a32 == 0).sum() > 1:
            continue                                  # ties after the f32 cast: order is the engine's own
        got = engine.topk_scores(a32, c["k"])
        want = [(float(np.float32(s)), i) for s, i in c["expected"]]
        assert got == want, (c["scores"], c["k"])
        n += 1
    assert n >= 50


@pytest.mark.parametrize("n,k", [(1, 1), (63, 10), (64, 64), (65, 7), (1000, 100), (4097, 1000), (100000, 100),
                                 (100000, 2048), (1 << 20, 100), (1 << 20, 1000), (3_000_000, 100)])
def test_random_vectors(engine, n, k):
    rng = np.random.default_rng(n * 31 + k)
    _check_exact(engine, rng.standard_normal(n), k)


def test_clustered_scores_like_the_1m_benchmark(engine):
    rng = np.random.default_rng(5)
    _check_exact(engine, 0.752 + 0.0072 * rng.standard_normal(1_000_000), 100)
    _check_exact(engine, 0.752 + 0.0072 * rng.standard_normal(1_000_000), 1000)


def test_ties_are_broken_by_ascending_index(engine):
    got = engine.topk_scores(np.ones(1000, dtype=np.float32), 10)
    assert got == [(1.0, i) for i in range(10)]                   # all tied: lowest indices win, ascending
    a = np.zeros(5000, dtype=np.float32)
    a[[7, 4000, 123]] = 2.0
    a[[9, 11]] = 1.0
    assert [i for _, i in engine.topk_scores(a, 6)] == [7, 123, 4000, 9, 11, 0]
    rng = np.random.default_rng(1)
    _check_exact(engine, rng.integers(0, 50, size=200_000).astype(np.float32), 300)     # massive ties


@pytest.mark.parametrize("order", ["ascending", "descending"])
def test_adversarial_sorted_input(engine, order):
    a = np.arange(300_000, dtype=np.float32)
    if order == "descending":
        a = a[::-1].copy()
    _check_exact(engine, a, 100)
    _check_exact(engine, a, 2048)


def test_special_values(engine):
    a = np.array([-1.0, 3.0, -3.0, 1e30, -1e30, 1e-38, -1e-38, np.inf, -np.inf, 0.5], dtype=np.float32)
    got = _check_exact(engine, a, 10)
    assert got[0][1] == 7 and got[-1][1] == 8
    # NaN is selected first, as np.argpartition treats it as the largest value (util.py:202)
    b = np.array([0.1, np.nan, 0.9, 0.2], dtype=np.float32)
    assert [i for _, i in engine.topk_scores(b, 2)] == [1, 2]
    # +0.0 and -0.0 are equal scores; the engine orders +0 first, then by index
    z = np.array([-0.0, 0.0, -0.0, 0.0], dtype=np.float32)
    assert sorted(i for _, i in engine.topk_scores(z, 4)) == [0, 1, 2, 3]


@pytest.mark.parametrize("n,k", [(3000, 3000), (5000, 4096), (70_000, 2049), (300_000, 100_000), (10_548, 10_548)])
def test_large_k_full_sort_path(engine, n, k):
    rng = np.random.default_rng(k)
    _check_exact(engine, rng.standard_normal(n), k)


def test_selection_is_idempotent_and_workspace_is_clean(engine):
    rng = np.random.default_rng(9)
    a = rng.standard_normal(500_000).astype(np.float32)
    first = engine.topk_scores(a, 100)
    for _ in range(3):
        assert engine.topk_scores(a, 100) == first                # group maxima were reset each time
    b = -a
    _check_exact(engine, b, 100)                                  # smaller maxima than the previous run
