"""GPU: the load path (SQLite blobs -> pinned slabs -> device matrix) and end-to-end retrieve against
golden results recorded from the real reference's svs.KB (oracle/make_golden.py)."""
import os
import shutil
import sqlite3

import numpy as np
import pytest

from _util import GOLDEN, golden_json, golden_npz, oracle

pytestmark = pytest.mark.gpu


def test_kb_small_sqlite_loads_bit_exact_and_retrieves_like_the_reference(tmp_path):
    import svs_b200
    dst = tmp_path / "kb.sqlite"
    shutil.copy(os.path.join(GOLDEN, "kb_small.sqlite"), dst)
    conn = sqlite3.connect(str(dst))
    g = golden_npz("kb_small_matrix.npz")
    exp = golden_json("kb_small.json")
    with svs_b200.Engine() as e:
        m = svs_b200.load_from_connection(e, conn)
        assert m.shape == g["matrix"].shape
        rows, ids = e.read_rows(0, m.shape[0])
        assert rows.tobytes() == g["matrix"].tobytes()            # reference src/svs/kb.py:573-618, bit-exact
        assert (ids == g["emb_ids"]).all()
        dev, bad = e.norm_stats()
        assert bad == 0 and dev < 1e-6                            # stub vectors are unit norm
        emb_to_doc = dict(conn.execute("SELECT embedding, id FROM docs WHERE embedding IS NOT NULL;").fetchall())
        for case in exp["queries"]:
            q = oracle.query_vec_of(case["vector"])
            got = m.retrieve(q, case["n"])
            want = [(r["score"], next(e_ for e_, d_ in emb_to_doc.items() if d_ == r["doc_id"])) for r in case["results"]]
            oracle.compare_retrieval(got, want, oracle.scores_of(g["matrix"], q), g["emb_ids"])
            # no near-ties in this fixture: the document ranking is identical to the reference's
            assert [emb_to_doc[i] for _, i in got] == [r["doc_id"] for r in case["results"]]


def test_many_slabs_and_chunked_loads():
    import svs_b200
    n, d = 70_000, 1536                       # ~430 MB: spans > 10 pinned slabs of 32 MB
    m = oracle.synth_matrix_normal(n, d, 41)
    ids = np.arange(10, 10 + n, dtype=np.int64)
    with svs_b200.Engine() as e:
        chunks = [(m[a:a + 9973], ids[a:a + 9973]) for a in range(0, n, 9973)]
        e.load_chunks(n, d, chunks)
        for a in (0, 12_345, n - 1000):
            rows, rid = e.read_rows(a, 1000)
            assert rows.tobytes() == m[a:a + 1000].tobytes() and (rid == ids[a:a + 1000]).all()
        with pytest.raises(svs_b200.EngineError):             # fewer rows than announced (kb.py:616)
            e.load_begin(10, 4)
            e.load_rows(np.zeros((3, 4), np.float32), np.arange(3))
            e.load_end()
        assert e.shape == (n, d)                               # failed load left the resident matrix alone
        with pytest.raises(svs_b200.EngineError):             # more rows than announced
            e.load_begin(2, 4)
            e.load_rows(np.zeros((3, 4), np.float32), np.arange(3))
        e.load_abort()
        q = m[777].copy()
        assert e.retrieve(q, 1)[0][1] == 787


def test_synthetic_generator_matches_the_oracle_bit_for_bit():
    import svs_b200
    with svs_b200.Engine() as e:
        e.load_synthetic(5000, 96, seed=7, id0=100, id_step=3)
        rows, ids = e.read_rows(1000, 64)
        raw = oracle.counter_uniform_rows(7, 1000, 64, 96)
        want = raw / np.sqrt((raw.astype(np.float64) ** 2).sum(axis=1))[:, None]
        np.testing.assert_allclose(rows, want, rtol=3e-7)        # same bits up to the norm's rounding
        assert (ids == 100 + 3 * np.arange(1000, 1064)).all()
        dev, bad = e.norm_stats()
        assert bad == 0 and dev < 1e-6


@pytest.mark.parametrize("devices", [[0], [0, 0, 0]])
def test_native_sqlite_scan_loads_the_same_device_matrix(tmp_path, devices, monkeypatch):
    """svsb_load_sqlite (parallel read-only libsqlite3 connections -> pinned slabs -> device rows) against the generic
    scan through a Python connection and against the oracle's build_embeddings_matrix (src/svs/kb.py:573-618)."""
    import svs_b200
    monkeypatch.setenv("SVSB_ALLOW_DUP_DEVICES", "1")
    g = golden_npz("kb_small_matrix.npz")
    with svs_b200.Engine(devices) as e:
        n, d = e.load_sqlite(os.path.join(GOLDEN, "kb_small.sqlite"))
        assert (n, d) == g["matrix"].shape == e.shape
        rows, ids = e.read_rows(0, n)
        assert rows.tobytes() == g["matrix"].tobytes() and (ids == g["emb_ids"]).all()
    # a larger table with holes in the rowids, d not a multiple of 4 (padded leading dimension), several threads
    path = str(tmp_path / "big.sqlite")
    conn = sqlite3.connect(path)
    conn.execute("CREATE TABLE embeddings (id INTEGER PRIMARY KEY, embedding BLOB NOT NULL);")
    rng = np.random.default_rng(5)
    src = rng.standard_normal((60_000, 130)).astype("<f4")
    src /= np.sqrt((src * src).sum(axis=1))[:, None]
    conn.executemany("INSERT INTO embeddings (embedding) VALUES (?);", ((r.tobytes(),) for r in src))
    conn.execute("DELETE FROM embeddings WHERE id % 11 = 3;")
    conn.commit()
    want_m, want_ids = oracle.build_embeddings_matrix(conn)
    q = oracle.synth_queries(2, 130, 6, dist="normal")
    with svs_b200.Engine(devices) as e, svs_b200.Engine([0]) as ref:
        for threads in (1, 3, 8):
            assert e.load_sqlite(path, threads=threads) == want_m.shape
            rows, ids = e.read_rows(0, len(want_ids))
            assert rows.tobytes() == want_m.tobytes() and (ids == want_ids).all()
        dev, bad = e.norm_stats()
        assert bad == 0 and dev < 1e-6
        m1 = svs_b200.load_from_connection(ref, conn)              # the generic path
        m2 = svs_b200.load_from_connection(e, conn, path=path)     # the native path through the same entry point
        for x in q:
            assert m1.retrieve(x, 50) == m2.retrieve(x, 50)
        oracle.compare_retrieval(m2.retrieve(q[0], 50), oracle.superheavy(want_m, want_ids, q[0], 50), oracle.scores_of(want_m, q[0]), want_ids)
        # what the native scan refuses is a refusal (SVSB_E_STATE), and load_from_connection falls back from it
        conn.execute("INSERT INTO embeddings (embedding) VALUES (?);", (b"\x00" * 12,))
        conn.commit()
        with pytest.raises(svs_b200.EngineError):
            e.load_sqlite(path)
        assert e.shape == want_m.shape                             # the resident matrix was left alone
        with pytest.raises(AssertionError):                        # the generic scan asserts like the reference (kb.py:613)
            svs_b200.load_from_connection(e, conn, path=path)
