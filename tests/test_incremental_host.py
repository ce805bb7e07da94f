"""CPU: the mutation log and the incremental-update protocol of the drop-in, with the real
reference package (oracle/_ref) on a real SQLite file and a NumPy restatement of svsb_apply_mutations as the engine.
After EVERY operation of a long random sequence -- bulk adds, bulk deletes, rolled-back transactions, id recycling,
set_doc_embedding-style replace -- the cached matrix the drop-in would query must equal what the reference's own
build_embeddings_matrix (src/svs/kb.py:573-618) reads from the file: same rows, same ids, same order."""
import sqlite3
import sys

import numpy as np
import pytest

from _util import oracle, reference_import_path, stub_vector

import svs_b200
from svs_b200 import matrix as matrix_mod
from svs_b200.mutations import LoggingConnection, MutationLog

from test_host_logic import FakeEngine


# ---- MutationLog by itself ------------------------------------------------------------------------------------
def test_log_commit_rollback_and_recycled_ids():
    log = MutationLog()
    log.begin(); log.record_insert(11, b"a" * 8); log.record_insert(12, b"b" * 8); log.commit()
    log.begin(); log.record_insert(13, b"c" * 8); log.rollback()                    # never happened
    log.begin(); log.record_delete(12); log.record_insert(12, b"d" * 8); log.record_delete(3); log.commit()
    dels, ids, blobs = log.take()
    assert dels == [3] and ids == [11, 12] and blobs == [b"a" * 8, b"d" * 8]          # 12 was replaced before anyone looked
    assert log.take() == ([], [], [])
    # a pre-existing row deleted and its id handed out again: tombstone + append, in that order
    log.begin(); log.record_delete(10); log.record_insert(10, b"e" * 8); log.commit()
    assert log.take() == ([10], [10], [b"e" * 8])


def test_log_poisons_itself_instead_of_guessing():
    log = MutationLog(max_bytes=32)
    log.begin(); log.record_insert(1, b"x" * 40); log.commit()
    assert log.take() is None and log.take() == ([], [], [])                          # too many bytes: rebuild, then a fresh start
    log.begin(); log.record_delete(5); log.commit()
    log.begin(); log.record_delete(5); log.commit()                                   # the same pre-existing row twice: impossible
    assert log.take() is None
    log.record_insert(9, b"")                                                         # outside any transaction
    assert log.take() is None


def test_logging_connection_sees_exactly_the_two_statements():
    conn = sqlite3.connect(":memory:")
    conn.execute("CREATE TABLE embeddings (id INTEGER PRIMARY KEY, embedding BLOB NOT NULL);")
    conn.execute("CREATE TABLE docs (id INTEGER PRIMARY KEY, embedding INTEGER);")
    log = MutationLog()
    c = LoggingConnection(conn, log)
    log.begin()
    c.execute("\n  INSERT INTO embeddings (embedding)\n  VALUES (?);\n", (b"\x00" * 8,))       # the reference's layout (kb.py:555-559)
    c.execute("INSERT INTO embeddings (embedding) VALUES (?);", (b"\x01" * 8,))
    c.execute("UPDATE docs SET embedding = ? WHERE id = ?;", (1, 1))
    assert c.execute("SELECT COUNT(*) FROM embeddings;").fetchone()[0] == 2
    c.execute("\n DELETE FROM embeddings WHERE id = ?;\n", (1,))
    c.execute("DELETE FROM embeddings WHERE id = ?;", (77,))                          # no such row: nothing to log
    log.commit()
    assert log.take() == ([], [2], [b"\x01" * 8])
    log.begin(); c.execute("DELETE FROM embeddings;"); log.commit()                   # not one of the reference's statements
    assert log.take() is None
    assert c.total_changes >= 3                                                       # everything else is the connection's


# ---- the whole protocol on the real reference package ----------------------------------------------------------
@pytest.fixture
def svs_fake_engine(monkeypatch):
    made = []

    def factory(devices=None):
        e = FakeEngine(devices, slab_rows=64)
        made.append(e)
        return e
    monkeypatch.setattr(matrix_mod, "Engine", factory)
    monkeypatch.syspath_prepend(reference_import_path())
    for k in [k for k in sys.modules if k == "svs" or k.startswith("svs.")]:
        monkeypatch.delitem(sys.modules, k)
    import svs
    svs_b200.install(svs)
    yield svs, made
    svs_b200.uninstall()


def _device_state(engine):
    live = engine.live if engine.live is not None else np.ones(engine.n, bool)
    return engine.rows[live], engine.ids[live]


@pytest.mark.skipif(reference_import_path() is None, reason="oracle/_ref (byte-compiled reference) not built")
def test_random_adds_deletes_and_rollbacks_track_a_fresh_rebuild(svs_fake_engine, tmp_path):
    svs, engines = svs_fake_engine
    d = 8
    path = str(tmp_path / "inc.sqlite")

    async def embed(texts):
        return [stub_vector(t, d) for t in texts]
    kb = svs.KB(path, embed)
    rng = np.random.default_rng(7)
    alive = []                                                     # doc ids with an embedding
    serial = 0
    with kb.bulk_add_docs() as add_doc:
        for _ in range(400):
            alive.append(add_doc(f"doc {serial}")); serial += 1
    check = sqlite3.connect(path)

    def verify():
        kb.retrieve("anything", 3)                                 # brings the device cache up to date
        eng = engines[-1]
        want_m, want_ids = oracle.build_embeddings_matrix(check)
        got_m, got_ids = _device_state(eng)
        assert got_ids.tolist() == want_ids.tolist()
        assert got_m.tobytes() == want_m.tobytes()
    verify()
    ops = {"add": 0, "del": 0, "rollback": 0, "recycle": 0, "noemb": 0}
    for step in range(300):
        op = rng.choice(["add", "del", "rollback", "recycle", "noemb"], p=[0.35, 0.3, 0.15, 0.1, 0.1])
        ops[op] += 1
        if op == "add":
            with kb.bulk_add_docs() as add_doc:
                for _ in range(int(rng.integers(1, 4))):
                    alive.append(add_doc(f"doc {serial}")); serial += 1
        elif op == "del" and len(alive) > 5:
            with kb.bulk_del_docs() as del_doc:
                for _ in range(int(rng.integers(1, 3))):
                    del_doc(alive.pop(int(rng.integers(0, len(alive)))))
        elif op == "rollback":
            with pytest.raises(RuntimeError):
                with kb.bulk_add_docs() as add_doc:
                    add_doc(f"never {serial}"); serial += 1
                    raise RuntimeError("abort this transaction")
        elif op == "recycle" and len(alive) > 5:
            # delete the newest document, then add: SQLite hands the same embeddings rowid out again (tests/test_kb.py:1592-1595)
            with kb.bulk_del_docs() as del_doc:
                del_doc(alive.pop())
            with kb.bulk_add_docs() as add_doc:
                alive.append(add_doc(f"doc {serial}")); serial += 1
        elif op == "noemb":
            with kb.bulk_add_docs() as add_doc:
                add_doc(f"plain {serial}", no_embedding=True); serial += 1
        verify()
    assert min(ops.values()) > 5
    eng = engines[-1]
    applies = [c for c in eng.calls if c[0] == "apply"]
    begins = [c for c in eng.calls if c[0] == "begin"]
    stats = kb.embeddings_matrix.device.stats
    assert len(applies) > 150 and stats["incremental_updates"] == len(applies)
    assert len(begins) == stats["full_builds"] < 10                # rebuilds only when tombstones pile up (compaction)
    assert stats["incremental_fallbacks"] == 0
    # close frees the device matrix (invalidate alone keeps it, stale)
    kb.close()
    assert ("close",) in eng.calls


@pytest.mark.skipif(reference_import_path() is None, reason="oracle/_ref (byte-compiled reference) not built")
def test_engine_refusal_falls_back_to_the_full_rebuild(svs_fake_engine, tmp_path, monkeypatch):
    svs, engines = svs_fake_engine
    d = 8

    async def embed(texts):
        return [stub_vector(t, d) for t in texts]
    kb = svs.KB(str(tmp_path / "fb.sqlite"), embed)
    with kb.bulk_add_docs() as add_doc:
        for i in range(20):
            add_doc(f"doc {i}")
    kb.retrieve("x", 1)
    eng = engines[-1]
    # somebody else changed the file behind the log's back: the engine refuses the delete of a row it does not hold
    raw = sqlite3.connect(str(tmp_path / "fb.sqlite"))
    eng.live = np.ones(eng.n, bool); eng.live[4] = False
    with kb.bulk_del_docs() as del_doc:
        del_doc(5)                                                 # embeddings.id 5 == row 4: "not live" in the engine
    res = kb.retrieve("x", 50)
    assert len(res) == 19 and kb.embeddings_matrix.device.stats["incremental_fallbacks"] == 1
    want_m, want_ids = oracle.build_embeddings_matrix(raw)
    got_m, got_ids = _device_state(engines[-1])
    assert got_ids.tolist() == want_ids.tolist() and got_m.tobytes() == want_m.tobytes()
    kb.close()
