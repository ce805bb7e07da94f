"""Pin the oracle (oracle/svs_oracle.py) against the reference: golden vectors produced by the real
reference (oracle/make_golden.py) and the reference's own known-answer tests, restated."""
import itertools
import os
import shutil
import sqlite3

import numpy as np
import pytest

from _util import GOLDEN, golden_json, golden_npz, oracle


# ---- get_top_k: reference tests/test_util.py:142-400 ------------------------------------------
def test_get_top_k_known_answers():
    # tests/test_util.py:143-149 (empty), 151-166 (one element), 168-197 ...
    assert oracle.get_top_k(np.array([]), 0) == []
    assert oracle.get_top_k(np.array([]), 1) == []
    assert oracle.get_top_k(np.array([0.4]), 0) == []
    assert oracle.get_top_k(np.array([0.4]), 1) == [(0.4, 0)]
    assert oracle.get_top_k(np.array([0.4]), 2) == [(0.4, 0)]                 # k clipped to len
    assert oracle.get_top_k(np.array([0.4, 0.2]), 1) == [(0.4, 0)]
    assert oracle.get_top_k(np.array([0.4, 0.2]), 2) == [(0.4, 0), (0.2, 1)]
    assert oracle.get_top_k(np.array([0.4, 0.2]), 3) == [(0.4, 0), (0.2, 1)]
    assert oracle.get_top_k(np.array([0.2, 0.4]), 1) == [(0.4, 1)]
    assert oracle.get_top_k(np.array([0.2, 0.4]), 2) == [(0.4, 1), (0.2, 0)]


def test_get_top_k_all_permutations_of_three():
    # the reference enumerates every permutation of three distinct scores for k = 0..4
    vals = [0.4, 0.2, 0.9]
    for perm in itertools.permutations(vals):
        a = np.array(perm)
        for k in range(0, 5):
            want = sorted(((s, i) for i, s in enumerate(perm)), reverse=True)[:max(0, min(k, 3))]
            assert oracle.get_top_k(a, k) == want


def test_get_top_k_tie_order_is_descending_index():
    # src/svs/util.py:203 sorts (score, index) tuples in reverse: equal scores -> larger index first
    a = np.array([1.0, 1.0, 1.0, 1.0, 1.0], dtype=np.float32)
    got = oracle.get_top_k(a, 5)
    assert [i for _, i in got] == [4, 3, 2, 1, 0]


def test_get_top_k_golden():
    arrays = golden_npz("topk_cases.npz")
    cases = golden_json("topk_cases.json")
    assert len(cases) >= 60
    for c in cases:
        got = oracle.get_top_k(arrays[c["scores"]], c["k"])
        assert [[s, i] for s, i in got] == c["expected"], (c["scores"], c["k"])


# ---- blob codec: reference tests/test_embeddings.py:13-22 -------------------------------------
def test_codec_known_answers():
    assert oracle.embedding_to_bytes([]) == b''
    assert oracle.embedding_to_bytes([1.0]) == b'\x00\x00\x80?'
    assert oracle.embedding_to_bytes([1.0, 3.5]) == b'\x00\x00\x80?\x00\x00`@'
    assert oracle.embedding_from_bytes(b'') == []
    assert oracle.embedding_from_bytes(b'\x00\x00\x80?') == [1.0]
    assert oracle.embedding_from_bytes(b'\x00\x00\x80?\x00\x00`@') == [1.0, 3.5]


def test_codec_golden():
    for c in golden_json("codec.json"):
        assert oracle.embedding_to_bytes(c["vector"]).hex() == c["hex"]
        assert oracle.embedding_from_bytes(bytes.fromhex(c["hex"])) == c["roundtrip"]


def test_magnitude_guard():
    # reference tests/test_embeddings.py:52-77: tolerance 1e-3
    assert not oracle.magnitude_ok([[1.0, 0.1, 0.0]])
    assert oracle.magnitude_ok([[1.0, 0.01, 0.0]])
    assert oracle.magnitude_ok([[1.0, 0.0, 0.0]])


# ---- matrix build: reference tests/test_kb.py:753-808 -----------------------------------------
def _mini_db(tmp_path, blobs):
    conn = sqlite3.connect(str(tmp_path / "t.sqlite"))
    conn.execute("CREATE TABLE embeddings (id INTEGER PRIMARY KEY, embedding BLOB NOT NULL);")
    for b in blobs:
        conn.execute("INSERT INTO embeddings (embedding) VALUES (?);", (b,))
    return conn


def test_build_embeddings_matrix_known_answer(tmp_path):
    conn = _mini_db(tmp_path, [b'\x00\x00\x80?\x00\x00`@', b'\x00\x00\x00@\x00\x00`@',
                               b'\x00\x00\x00@\x00\x00\x80?', b'\x00\x00`@\x00\x00\x80@'])
    m, ids = oracle.build_embeddings_matrix(conn)
    assert (m == np.array([[1.0, 3.5], [2.0, 3.5], [2.0, 1.0], [3.5, 4.0]])).all()
    assert m.dtype == np.float32 and ids.dtype == np.int64
    assert (ids == np.array([1, 2, 3, 4])).all()
    conn.execute("DELETE FROM embeddings WHERE id = 3;")
    m, ids = oracle.build_embeddings_matrix(conn)
    assert (m == np.array([[1.0, 3.5], [2.0, 3.5], [3.5, 4.0]])).all()
    assert (ids == np.array([1, 2, 4])).all()


def test_build_embeddings_matrix_empty(tmp_path):
    m, ids = oracle.build_embeddings_matrix(_mini_db(tmp_path, []))
    assert m.shape == (0, 0) and ids.shape == (0,)                      # kb.py:595-601
    with pytest.raises(ValueError):
        oracle.scores_of(m, np.zeros(3, dtype=np.float32))              # np.dot (0,0).(3,) -> ValueError


def test_build_embeddings_matrix_golden(tmp_path):
    g = golden_npz("kb_small_matrix.npz")
    dst = tmp_path / "kb.sqlite"
    shutil.copy(os.path.join(GOLDEN, "kb_small.sqlite"), dst)
    conn = sqlite3.connect(str(dst))
    m, ids = oracle.build_embeddings_matrix(conn)
    assert m.tobytes() == g["matrix"].tobytes()                          # bit-exact
    assert (ids == g["emb_ids"]).all()
    assert (np.diff(ids) > 0).all()                                      # rowid scan: ascending ids


# ---- superheavy(): reference src/svs/kb.py:1622-1627 -------------------------------------------
@pytest.mark.parametrize("name", ["superheavy_d96.npz", "superheavy_d1536.npz"])
def test_superheavy_golden(name):
    g = golden_npz(name)
    m, ids, qs = g["matrix"], g["emb_ids"], g["queries"]
    for qi in range(len(qs)):
        x = oracle.scores_of(m, qs[qi])
        assert x.dtype == np.float32
        # same NumPy/BLAS build as the generator -> bit-identical; allow 1 ulp-ish drift otherwise
        np.testing.assert_allclose(x, g[f"scores_q{qi}"], rtol=2e-6, atol=1e-7)
        for k in g["ks"]:
            got = oracle.superheavy(m, ids, qs[qi], int(k))
            want = list(zip(g[f"top_q{qi}_k{k}_scores"].tolist(), g[f"top_q{qi}_k{k}_ids"].tolist()))
            rep = oracle.compare_retrieval(
                [(s, e) for s, e in sorted(got, key=lambda t: (-t[0], t[1]))],
                want, g[f"scores_q{qi}"], ids)
            assert rep["n"] == min(int(k), len(ids))


def test_kb_small_golden_against_oracle(tmp_path):
    """svs.KB.retrieve outputs recorded from the real reference == oracle on the same SQLite file."""
    exp = golden_json("kb_small.json")
    dst = tmp_path / "kb.sqlite"
    shutil.copy(os.path.join(GOLDEN, "kb_small.sqlite"), dst)
    conn = sqlite3.connect(str(dst))
    m, ids = oracle.build_embeddings_matrix(conn)
    emb_to_doc = dict(conn.execute("SELECT embedding, id FROM docs WHERE embedding IS NOT NULL;").fetchall())
    for case in exp["queries"]:
        q = oracle.query_vec_of(case["vector"])
        got = oracle.superheavy(m, ids, q, case["n"])
        assert len(got) == len(case["results"]) == min(case["n"], len(ids))
        for (s, e), r in zip(got, case["results"]):
            assert emb_to_doc[e] == r["doc_id"]
            assert s == pytest.approx(r["score"], rel=2e-6, abs=1e-7)


# ---- comparator and generators -----------------------------------------------------------------
def test_comparator_accepts_near_tie_swaps_only():
    ids = np.arange(10, 16, dtype=np.int64)
    x = np.array([0.9, 0.5, 0.5000001, 0.1, 0.7, 0.3], dtype=np.float32)
    orc = [(float(x[i]), int(ids[i])) for i in np.argsort(-x)][:3]
    ok = [(float(x[0]), 10), (float(x[4]), 14), (float(x[1]), 11)]       # 11 <-> 12 are a near tie
    oracle.compare_retrieval(ok, orc, x, ids)
    bad = [(float(x[0]), 10), (float(x[4]), 14), (float(x[5]), 15)]      # 0.3 is not near 0.5
    with pytest.raises(AssertionError):
        oracle.compare_retrieval(bad, orc, x, ids)
    wrong_order = [(float(x[4]), 14), (float(x[0]), 10), (float(x[2]), 12)]
    with pytest.raises(AssertionError):
        oracle.compare_retrieval(wrong_order, orc, x, ids)
    dup = [(float(x[0]), 10), (float(x[0]), 10), (float(x[2]), 12)]
    with pytest.raises(AssertionError):
        oracle.compare_retrieval(dup, orc, x, ids)


def test_canonical_top_k_order():
    s = np.array([0.5, 0.9, 0.5, 0.9, 0.1], dtype=np.float32)
    ids = np.array([7, 3, 2, 9, 1], dtype=np.int64)
    assert oracle.canonical_top_k(s, ids, 4) == [(float(s[1]), 3), (float(s[3]), 9), (0.5, 2), (0.5, 7)]
    assert oracle.canonical_top_k(s, ids, 0) == []
    assert len(oracle.canonical_top_k(s, ids, 99)) == 5


def test_counter_generator_is_a_pure_function_of_seed_row_col():
    a = oracle.counter_uniform_rows(3, 0, 64, 48)
    b = oracle.counter_uniform_rows(3, 32, 32, 48)
    assert a.dtype == np.float32 and a.shape == (64, 48)
    assert (a[32:] == b).all()
    assert 0.0 <= a.min() and a.max() < 1.0
    assert abs(float(a.mean()) - 0.5) < 0.02
    assert not (a == oracle.counter_uniform_rows(4, 0, 64, 48)).all()
    # pin the hash itself with an independent big-int restatement (the CUDA generator must match)
    M64 = (1 << 64) - 1

    def mix(z):
        z ^= z >> 33; z = (z * 0xFF51AFD7ED558CCD) & M64
        z ^= z >> 33; z = (z * 0xC4CEB9FE1A85EC53) & M64
        z ^= z >> 33
        return z

    for seed, row, col, d in [(0, 0, 0, 2), (0, 0, 1, 2), (3, 40, 7, 48), (2**40 + 5, 9_999_999, 1535, 1536)]:
        ctr = ((row * d + col) * 0x9E3779B97F4A7C15 + seed * 0xC4CEB9FE1A85EC53 + 1) & M64
        want = np.float32(mix(ctr) >> 40) * np.float32(2.0 ** -24)
        got = oracle.counter_uniform_rows(seed, row, 1, d)[0, col]
        assert got == want, (seed, row, col)


def test_synth_matrices_are_unit_norm():
    for f in (oracle.synth_matrix_uniform, oracle.synth_matrix_normal):
        m = f(200, 64, 0)
        assert m.dtype == np.float32
        assert oracle.magnitude_ok(m.tolist(), 1e-5)


def test_get_top_pairs_known_answers():
    """The known-answer cases of the reference's tests/test_util.py:403-470 (get_top_pairs)."""
    with pytest.raises(AssertionError):
        oracle.get_top_pairs(np.zeros((3, 2, 5)), top_k=3)           # not 2-D
    with pytest.raises(AssertionError):
        oracle.get_top_pairs(np.zeros((3, 2)), top_k=3)              # not square
    assert oracle.get_top_pairs(np.zeros((0, 0)), top_k=3) == []
    assert oracle.get_top_pairs(np.array([[1]]), top_k=3) == []
    assert oracle.get_top_pairs(np.array([[1, 2], [9, 4]]), top_k=3) == [(2, 0, 1)]
    assert oracle.get_top_pairs(np.array([[1, 2, 3], [9, 5, 6], [9, 9, 9]]), top_k=3) == [(6, 1, 2), (3, 0, 2), (2, 0, 1)]
    m = np.array([[1, 2, 3, 4], [20, 20, 7, 8], [20, 20, 20, 12], [20, 20, 20, 16]])
    assert oracle.get_top_pairs(m, top_k=3) == [(12, 2, 3), (8, 1, 3), (7, 1, 2)]


def test_top_pairwise_and_its_comparator():
    rng = np.random.default_rng(21)
    m = oracle.synth_matrix_normal(60, 12, 2)
    ids = np.cumsum(rng.integers(1, 4, size=60)).astype(np.int64)
    pairwise = np.dot(m, m.T)
    want = oracle.top_pairwise(m, ids, 40)
    assert len(want) == 40 and all(want[i][0] >= want[i + 1][0] for i in range(39))
    oracle.compare_pairs(want, want, pairwise, ids)
    wrong = list(want)
    wrong[3] = (wrong[3][0], wrong[30][1], wrong[30][2])             # a far-away pair smuggled into rank 3
    with pytest.raises(AssertionError):
        oracle.compare_pairs(wrong, want, pairwise, ids)
