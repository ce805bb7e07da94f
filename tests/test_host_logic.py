"""Host-side logic of the drop-in (no GPU): the SQLite scan -> staging slabs, the cache/invalidate
protocol of DeviceEmbeddingsMatrix, and install()/uninstall() on a stand-in `svs` module.
A recording fake replaces the Engine; numerics are NOT under test here."""
import asyncio
import os
import shutil
import sqlite3
import types

import numpy as np
import pytest

from _util import GOLDEN, golden_npz, reference_import_path

import svs_b200
from svs_b200 import matrix as matrix_mod


class FakeSnapshot:
    batches = []                                   # sizes of the retrieve_many calls seen (coalescer test)

    def __init__(self, eng):
        live = eng.live if eng.live is not None else np.ones(eng.n, bool)
        self.shape = (int(live.sum()), eng.d if eng.n else 0)
        self.generation = eng.gen
        self.physical_rows, self.live_rows = eng.n, int(live.sum())
        self._rows = eng.rows[live].copy()
        self._ids = eng.ids[live].copy()

    def retrieve(self, q, n):
        return [(0.0, int(i)) for i in self._ids[:max(0, n)]]

    def retrieve_many(self, qs, n):
        FakeSnapshot.batches.append(len(qs))
        return [self.retrieve(q, n) for q in qs]

    def top_pairs(self, n):
        return [(0.5, int(self._ids[0]), int(self._ids[1]))][:max(0, n)] if len(self._ids) > 1 else []


class FakeEngine:
    """Records what the host logic asks of the C ABI; slab capacity is tiny to force many slabs."""

    def __init__(self, devices=None, slab_rows=7):
        self.slab_rows = slab_rows
        self.gen = 0
        self.loaded = False
        self.calls = []
        self.n = self.d = 0
        self.rows = np.zeros((0, 0), np.float32)
        self.ids = np.zeros(0, np.int64)
        self.live = None

    def load_begin(self, n, d, normalize=False):
        self.calls.append(("begin", n, d, normalize))
        self._n, self._d = n, d
        self._rows = np.zeros((n, d), np.float32)
        self._ids = np.zeros(n, np.int64)
        self._filled = 0

    def acquire_slab(self, d):
        cap = min(self.slab_rows, self._n - self._filled)
        self._slab = np.zeros(cap * d * 4, np.uint8)
        self._slab_ids = np.zeros(cap, np.int64)
        return self._slab, self._slab_ids

    def commit_slab(self, count):
        self.calls.append(("commit", count))
        if count:
            self._rows[self._filled:self._filled + count] = self._slab[:count * self._d * 4].view(np.float32).reshape(count, self._d)
            self._ids[self._filled:self._filled + count] = self._slab_ids[:count]
            self._filled += count

    def load_rows(self, rows, ids):
        self._ids[self._filled:self._filled + len(ids)] = ids
        self._filled += len(ids)

    def load_end(self):
        assert self._filled == self._n
        self.n, self.d, self.rows, self.ids, self.live = self._n, self._d, self._rows, self._ids, None
        self.gen += 1
        self.loaded = True
        self.calls.append(("end",))
        return self.gen

    def load_abort(self):
        self.calls.append(("abort",))

    def invalidate(self):
        self.calls.append(("invalidate",))
        self.loaded = False

    def generation_rows(self):
        live = self.live if self.live is not None else np.ones(self.n, bool)
        return self.n, int(live.sum())

    def apply_mutations(self, del_ids, add_ids, add_rows):
        """The contract of svsb_apply_mutations (include/svsb200.h) restated in NumPy."""
        import svs_b200
        from svs_b200 import _lib
        self.calls.append(("apply", list(del_ids), list(add_ids)))
        def refuse(why):
            raise svs_b200.EngineError(_lib.SVSB_E_STATE, why)
        if not self.loaded or self.n == 0:
            refuse("nothing resident")
        live = self.live.copy() if self.live is not None else np.ones(self.n, bool)
        if len(set(del_ids)) != len(del_ids):
            refuse("duplicate delete")
        for e in del_ids:
            hit = np.nonzero((self.ids == e) & live)[0]
            if len(hit) != 1:
                refuse(f"id {e} is not live")
            live[hit[0]] = False
        add_ids = list(add_ids)
        if add_ids:
            if add_rows.shape != (len(add_ids), self.d) or any(b <= a for a, b in zip(add_ids, add_ids[1:])):
                refuse("bad rows")
            if live.any() and add_ids[0] <= self.ids[live].max():
                refuse("id not above every live id")
            self.rows = np.concatenate([self.rows, np.asarray(add_rows, np.float32)])
            self.ids = np.concatenate([self.ids, np.asarray(add_ids, np.int64)])
            live = np.concatenate([live, np.ones(len(add_ids), bool)])
        self.n, self.live = len(self.ids), live
        self.gen += 1
        return self.gen

    def snapshot(self):
        return FakeSnapshot(self)

    def close(self):
        self.calls.append(("close",))


@pytest.fixture
def fake_engine(monkeypatch):
    made = []

    def factory(devices=None):
        e = FakeEngine(devices)
        made.append(e)
        return e
    monkeypatch.setattr(matrix_mod, "Engine", factory)
    return made


def _kb_copy(tmp_path):
    dst = tmp_path / "kb.sqlite"
    shutil.copy(os.path.join(GOLDEN, "kb_small.sqlite"), dst)
    return sqlite3.connect(str(dst), check_same_thread=False)   # as svs.kb._DB does (kb.py:782)


def test_load_from_connection_streams_the_scan_into_slabs(tmp_path):
    g = golden_npz("kb_small_matrix.npz")
    eng = FakeEngine(slab_rows=7)
    m = svs_b200.load_from_connection(eng, _kb_copy(tmp_path))
    assert m.shape == g["matrix"].shape == (415, 64)
    assert eng.rows.tobytes() == g["matrix"].tobytes()           # bit-exact, scan order
    assert (eng.ids == g["emb_ids"]).all()
    commits = [c[1] for c in eng.calls if c[0] == "commit"]
    assert commits[:-1] == [7] * 59 + [2] and commits[-1] == 0   # 415 rows through 7-row slabs
    assert eng.calls[0] == ("begin", 415, 64, False) and eng.calls[-1] == ("end",)


def test_load_from_connection_empty_table(tmp_path):
    conn = sqlite3.connect(str(tmp_path / "e.sqlite"))
    conn.execute("CREATE TABLE embeddings (id INTEGER PRIMARY KEY, embedding BLOB NOT NULL);")
    eng = FakeEngine()
    m = svs_b200.load_from_connection(eng, conn)
    assert m.shape == (0, 0)                                      # kb.py:595-601
    assert eng.calls == [("begin", 0, 0, False), ("end",)]


def test_load_from_connection_rejects_ragged_rows(tmp_path):
    conn = sqlite3.connect(str(tmp_path / "r.sqlite"))
    conn.execute("CREATE TABLE embeddings (id INTEGER PRIMARY KEY, embedding BLOB NOT NULL);")
    conn.execute("INSERT INTO embeddings (embedding) VALUES (?);", (b"\x00" * 8,))
    conn.execute("INSERT INTO embeddings (embedding) VALUES (?);", (b"\x00" * 12,))
    eng = FakeEngine()
    with pytest.raises(AssertionError):                           # the reference asserts too (kb.py:613)
        svs_b200.load_from_connection(eng, conn)
    assert ("abort",) in eng.calls and ("end",) not in eng.calls


class _FakeDB:
    """`with db as q:` yields an object with .conn, like svs.kb._DB / _Querier."""

    def __init__(self, conn):
        self.conn = conn
        self.entered = 0

    def __enter__(self):
        self.entered += 1
        return types.SimpleNamespace(conn=self.conn)

    def __exit__(self, *a):
        return False


def test_device_embeddings_matrix_cache_protocol(tmp_path, fake_engine):
    db = _FakeDB(_kb_copy(tmp_path))
    cache = svs_b200.DeviceEmbeddingsMatrix()
    assert fake_engine == []                                      # engine is created lazily
    m1 = cache.get_sync(db)
    assert db.entered == 1 and len(fake_engine) == 1
    assert cache.get_sync(db) is m1 and db.entered == 1           # hit: no rebuild (kb.py:867-869)
    cache.invalidate()                                            # marks the device matrix stale, keeps it resident ...
    assert ("invalidate",) not in fake_engine[0].calls
    m2 = cache.get_sync(db)                                       # ... no mutation log on this db: the full rebuild, old one dropped first
    assert ("invalidate",) in fake_engine[0].calls
    assert m2 is not m1 and db.entered == 2 and m2.generation == m1.generation + 1
    # the old handle still answers from the generation it pinned
    assert m1.retrieve(np.zeros(64, np.float32), 3) == m2.retrieve(np.zeros(64, np.float32), 3)

    async def go():
        cache.invalidate()
        m3 = await cache.get(db)
        assert (await cache.get(db)) is m3
        return m3
    m3 = asyncio.run(go())
    assert m3.shape == (415, 64)
    cache.close()
    assert ("close",) in fake_engine[0].calls


def test_prewarm_creates_the_engine_in_the_background_and_loads_wait_for_it(tmp_path, fake_engine):
    db = _FakeDB(_kb_copy(tmp_path))
    cache = svs_b200.DeviceEmbeddingsMatrix()
    t = cache.prewarm()
    m = cache.get_sync(db)                                        # may race with the pre-warm thread: one engine either way
    t.join(timeout=10)
    assert not t.is_alive() and len(fake_engine) == 1 and m.shape == (415, 64)
    cache.close()


def _fake_svs_module():
    """A stand-in with the seam of svs.kb (reference src/svs/kb.py:856-893, 925+, 1407+)."""
    kb = types.ModuleType("fakesvs.kb")

    class _EmbeddingsMatrix:
        def __init__(self):
            self.embeddings_matrix = None
            self.emb_id_lookup = None
            self.invalidations = 0

        def invalidate(self):
            self.invalidations += 1
            self.embeddings_matrix = None
            self.emb_id_lookup = None

    class KB:
        def __init__(self):
            self.embeddings_matrix = kb._EmbeddingsMatrix()

        def retrieve(self, query, n):
            return "host"

        def document_top_pairwise_scores(self, n):
            return "host pairs"

    class AsyncKB(KB):
        async def retrieve(self, query, n):
            return "host"

        async def load(self):
            return "host"

        async def document_top_pairwise_scores(self, n):
            return "host pairs"

    import logging
    kb._EmbeddingsMatrix, kb.KB, kb.AsyncKB, kb._LOG = _EmbeddingsMatrix, KB, AsyncKB, logging.getLogger("fakesvs")
    mod = types.ModuleType("fakesvs")
    mod.kb = kb
    return mod


def test_install_and_uninstall_patch_only_the_seam(fake_engine):
    mod = _fake_svs_module()
    orig_cls, orig_retrieve = mod.kb._EmbeddingsMatrix, mod.kb.KB.retrieve
    svs_b200.install(mod)
    try:
        assert mod.kb._EmbeddingsMatrix is not orig_cls and issubclass(mod.kb._EmbeddingsMatrix, orig_cls)
        assert mod.kb.KB.retrieve is not orig_retrieve
        kb = mod.kb.KB()
        assert isinstance(kb.embeddings_matrix.device, svs_b200.DeviceEmbeddingsMatrix)
        kb.embeddings_matrix.invalidate()                          # drops the host cache, marks the device cache stale
        assert kb.embeddings_matrix.invalidations == 1
        svs_b200.install(mod)                                      # idempotent
    finally:
        svs_b200.uninstall()
    assert mod.kb._EmbeddingsMatrix is orig_cls and mod.kb.KB.retrieve is orig_retrieve
    assert mod.kb.KB().retrieve("q", 1) == "host"
    assert mod.kb.KB().document_top_pairwise_scores(1) == "host pairs"
    assert not hasattr(mod.kb.KB, "retrieve_many") and not hasattr(mod.kb.AsyncKB, "retrieve_many")


@pytest.mark.skipif(reference_import_path() is None, reason="oracle/_ref (byte-compiled reference) not built")
def test_install_on_the_real_reference_package(fake_engine, tmp_path, monkeypatch):
    """With the real svs package: a patched KB loads through the (fake) engine and returns documents."""
    import sys
    monkeypatch.syspath_prepend(reference_import_path())
    for k in [k for k in sys.modules if k == "svs" or k.startswith("svs.")]:
        monkeypatch.delitem(sys.modules, k)
    import svs
    svs_b200.install(svs)
    try:
        async def embed(texts):
            return [[1.0, 0.0, 0.0] for _ in texts]
        kb = svs.KB(str(tmp_path / "x.sqlite"), embed)
        with kb.bulk_add_docs() as add_doc:
            for t in ("a", "b", "c"):
                add_doc(t)
        res = kb.retrieve("q", 2)                                  # FakeSnapshot returns the first ids
        assert [r["doc"]["text"] for r in res] == ["a", "b"]
        assert fake_engine[0].calls[0] == ("begin", 3, 3, False)
        with kb.bulk_del_docs() as del_doc:
            del_doc(1)
        assert [r["doc"]["text"] for r in kb.retrieve("q", 5)] == ["b", "c"]      # kb.py:1541 reached the device cache ...
        assert ("apply", [1], []) in fake_engine[0].calls         # ... as an incremental update: one tombstone, no rebuild
        assert [c[0] for c in fake_engine[0].calls].count("begin") == 1
        pair = kb.document_top_pairwise_scores(1)[0]               # pairwise now asks the (fake) engine for the pair
        assert (pair[0], pair[1]["text"], pair[2]["text"]) == (0.5, "b", "c")
        many = kb.retrieve_many(["q1", "q2", "q3"], 1)             # additive batched API: one embed call, one engine call
        assert [[r["doc"]["text"] for r in res] for res in many] == [["b"], ["b"], ["b"]]
        assert kb.retrieve_many([], 3) == []
        kb.close()
    finally:
        svs_b200.uninstall()


@pytest.mark.skipif(reference_import_path() is None, reason="oracle/_ref (byte-compiled reference) not built")
def test_concurrent_async_retrieves_are_coalesced_into_batches(fake_engine, tmp_path, monkeypatch):
    """AsyncKB.retrieve calls that are in flight together reach the engine as retrieve_many batches; every caller
    still gets its own n results; a lone call goes out as a single query."""
    import sys
    monkeypatch.syspath_prepend(reference_import_path())
    for k in [k for k in sys.modules if k == "svs" or k.startswith("svs.")]:
        monkeypatch.delitem(sys.modules, k)
    import svs
    svs_b200.install(svs)
    try:
        async def embed(texts):
            return [[1.0, 0.0, 0.0] for _ in texts]

        async def go():
            kb = svs.AsyncKB(str(tmp_path / "a.sqlite"), embed)
            async with kb.bulk_add_docs() as add_doc:
                for t in ("a", "b", "c", "d"):
                    await add_doc(t)
            FakeSnapshot.batches.clear()
            one = await kb.retrieve("q", 2)
            assert [r["doc"]["text"] for r in one] == ["a", "b"] and FakeSnapshot.batches == []
            outs = await asyncio.gather(*[kb.retrieve(f"q{i}", 1 + i % 3) for i in range(12)])
            assert [len(o) for o in outs] == [1 + i % 3 for i in range(12)]
            assert all(o[0]["doc"]["text"] == "a" for o in outs)
            assert sum(FakeSnapshot.batches) >= 8 and max(FakeSnapshot.batches) > 1       # most of them travelled together
            await kb.close()
        asyncio.run(go())
    finally:
        svs_b200.uninstall()
