"""CPU: the native SQLite scan (svs_b200/csrc/loader.cu -- libsqlite3 bound at run time, parallel read-only connections
over contiguous rowid ranges) against the oracle's build_embeddings_matrix (reference src/svs/kb.py:573-618) on the same
file: same rows, same ids, same order, bit for bit, for any thread count; the reference's asserts become errors the
caller can fall back from."""
import os
import sqlite3

import numpy as np
import pytest

from _util import GOLDEN, golden_npz, oracle

import svs_b200
from svs_b200 import _lib
from svs_b200.engine import sqlite_read


def _make_db(path, n, d, seed=0, delete_every=0, recycle=False):
    rng = np.random.default_rng(seed)
    conn = sqlite3.connect(path)
    conn.execute("CREATE TABLE embeddings (id INTEGER PRIMARY KEY, embedding BLOB NOT NULL);")    # kb.py:80-83
    rows = rng.standard_normal((n, d)).astype("<f4")
    conn.executemany("INSERT INTO embeddings (embedding) VALUES (?);", ((r.tobytes(),) for r in rows))
    if delete_every:
        conn.execute("DELETE FROM embeddings WHERE id % ? = 0;", (delete_every,))
    if recycle:                                                   # the newest ids deleted and handed out again (different blobs)
        conn.execute("DELETE FROM embeddings WHERE id > ?;", (n - 50,))
        extra = rng.standard_normal((80, d)).astype("<f4")
        conn.executemany("INSERT INTO embeddings (embedding) VALUES (?);", ((r.tobytes(),) for r in extra))
    conn.commit()
    return conn


def test_library_binds_libsqlite3():
    assert _lib.load().svsb_sqlite_available() == 1


@pytest.mark.parametrize("threads", [1, 2, 5, 8])
def test_scan_equals_the_oracle_for_any_thread_count(tmp_path, threads):
    path = str(tmp_path / "t.sqlite")
    conn = _make_db(path, 40_000, 24, seed=threads, delete_every=7, recycle=True)
    want_m, want_ids = oracle.build_embeddings_matrix(conn)
    m, ids = sqlite_read(path, threads)
    assert ids.tolist() == want_ids.tolist()
    assert m.tobytes() == want_m.tobytes()
    assert (np.diff(ids) > 0).all()                               # rowid order == scan order (kb.py:603-609)


def test_golden_kb_and_sparse_rowids(tmp_path):
    g = golden_npz("kb_small_matrix.npz")
    m, ids = sqlite_read(os.path.join(GOLDEN, "kb_small.sqlite"), 4)
    assert m.tobytes() == g["matrix"].tobytes() and (ids == g["emb_ids"]).all()
    # rowids spread over a huge range: the ranges are split by VALUE, most are empty, the order must still hold
    path = str(tmp_path / "sparse.sqlite")
    conn = sqlite3.connect(path)
    conn.execute("CREATE TABLE embeddings (id INTEGER PRIMARY KEY, embedding BLOB NOT NULL);")
    rng = np.random.default_rng(3)
    rid = np.unique(rng.integers(-2**40, 2**62, size=6000)).astype(np.int64)
    rows = rng.standard_normal((len(rid), 8)).astype("<f4")
    conn.executemany("INSERT INTO embeddings (id, embedding) VALUES (?, ?);", ((int(i), r.tobytes()) for i, r in zip(rid, rows)))
    conn.commit()
    m, ids = sqlite_read(path, 6)
    assert ids.tolist() == rid.tolist() and m.tobytes() == rows.tobytes()


def test_empty_table_ragged_rows_and_missing_file(tmp_path):
    path = str(tmp_path / "e.sqlite")
    conn = sqlite3.connect(path)
    conn.execute("CREATE TABLE embeddings (id INTEGER PRIMARY KEY, embedding BLOB NOT NULL);")
    conn.commit()
    m, ids = sqlite_read(path)
    assert m.shape == (0, 0) and ids.shape == (0,)                # kb.py:595-601
    conn.execute("INSERT INTO embeddings (embedding) VALUES (?);", (b"\x00" * 8,))
    conn.execute("INSERT INTO embeddings (embedding) VALUES (?);", (b"\x00" * 12,))
    conn.commit()
    with pytest.raises(svs_b200.EngineError, match="unequal length") as ex:      # the reference asserts (kb.py:613)
        sqlite_read(path)
    assert ex.value.code == _lib.SVSB_E_STATE
    with pytest.raises(svs_b200.EngineError, match="cannot open"):
        sqlite_read(str(tmp_path / "nope" / "missing.sqlite"))
    conn.execute("DELETE FROM embeddings;")
    conn.execute("INSERT INTO embeddings (embedding) VALUES (?);", (b"\x00" * 7,))
    conn.commit()
    with pytest.raises(svs_b200.EngineError, match="multiple of 4"):             # embeddings/util.py:20-21
        sqlite_read(path)


def test_uncommitted_changes_of_another_connection_are_not_seen(tmp_path):
    """The scan runs on its own read-only connections: it sees the committed state, like any second reader."""
    path = str(tmp_path / "u.sqlite")
    conn = _make_db(path, 300, 4)
    conn.execute("INSERT INTO embeddings (embedding) VALUES (?);", (b"\x00" * 16,))         # not committed
    m, ids = sqlite_read(path, 2)
    assert len(ids) == 300
    conn.commit()
    assert len(sqlite_read(path, 2)[1]) == 301


def test_row_counts_come_from_the_docs_index_and_are_verified(tmp_path, monkeypatch):
    """With the reference's docs table present the per-range counts come from idx_docs_embedding (kb.py:96); the scan checks
    them (kb.py:616), so an embeddings row no document points at is a refusal (the caller falls back), never a wrong matrix."""
    path = str(tmp_path / "d.sqlite")
    conn = _make_db(path, 9000, 8, seed=2)
    conn.execute("CREATE TABLE docs (id INTEGER PRIMARY KEY, embedding INTEGER REFERENCES embeddings(id));")
    conn.execute("CREATE INDEX idx_docs_embedding ON docs(embedding);")
    conn.execute("INSERT INTO docs (embedding) SELECT id FROM embeddings;")
    conn.execute("INSERT INTO docs (embedding) VALUES (NULL);")                   # a document without an embedding
    conn.commit()
    want_m, want_ids = oracle.build_embeddings_matrix(conn)
    for threads in (1, 2):
        m, ids = sqlite_read(path, threads)
        assert ids.tolist() == want_ids.tolist() and m.tobytes() == want_m.tobytes()
    conn.execute("DELETE FROM docs WHERE embedding = 4000;")                       # orphan: embeddings row 4000 stays
    conn.commit()
    with pytest.raises(svs_b200.EngineError, match="COUNT") as ex:
        sqlite_read(path, 2)
    assert ex.value.code == _lib.SVSB_E_STATE
    monkeypatch.setenv("SVSB_LOAD_COUNT_VIA_DOCS", "0")                            # counting in the table itself is always right
    m, ids = sqlite_read(path, 2)
    assert ids.tolist() == want_ids.tolist()
