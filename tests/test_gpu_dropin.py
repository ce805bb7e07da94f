"""GPU: the reference's own svs.KB / svs.AsyncKB (byte-compiled into oracle/_ref by oracle/build_ref.py)
with svs_b200.install() applied -- the drop-in boundary end to end.  Mirrors the reference's
tests/test_kb.py:1738-1848 (sync) and 1204-1318 (async)."""
import asyncio
import os
import sys

import numpy as np
import pytest

from _util import oracle, reference_import_path, stub_vector

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(reference_import_path() is None, reason="oracle/_ref (byte-compiled reference) not built")]


@pytest.fixture
def svs_patched(monkeypatch):
    monkeypatch.syspath_prepend(reference_import_path())
    for k in [k for k in sys.modules if k == "svs" or k.startswith("svs.")]:
        monkeypatch.delitem(sys.modules, k)
    import svs
    import svs_b200
    svs_b200.install(svs)
    yield svs
    svs_b200.uninstall()


async def _embed_words(texts):
    ret = []
    for text in texts:
        if 'first' in text:
            ret.append([1.0, 0.001, 0.0])
        elif 'second' in text:
            ret.append([0.0, 1.0, 0.0001])
        elif 'third' in text:
            ret.append([0.01, 0.0, 1.0])
        elif 'forth' in text:
            ret.append([0.707, 0.707, 0.0])
        else:
            raise ValueError("unexpected doc")
    return ret


def test_kb_retrieve_et_al_sync(svs_patched, tmp_path):
    svs = svs_patched
    path = str(tmp_path / "testdb.sqlite")
    kb = svs.KB(path, _embed_words)
    with kb.bulk_add_docs() as add_doc:
        assert add_doc("third doc") == 1
        assert add_doc("first doc") == 2
        assert add_doc("second doc") == 3
    kb.close()

    kb = svs.KB(path, _embed_words)
    texts = lambda q, n: [d['doc']['text'] for d in kb.retrieve(q, n=n)]
    assert texts('... first ...', 3) == ['first doc', 'third doc', 'second doc']
    assert texts('... second ...', 3) == ['second doc', 'first doc', 'third doc']
    assert texts('... third ...', 3) == ['third doc', 'first doc', 'second doc']
    res = kb.retrieve('... first ...', 3)
    assert isinstance(res[0]['score'], float) and res[0]['score'] == pytest.approx(1.000001, rel=1e-6)
    assert res[0]['doc']['embedding'] is True
    # the untouched pairwise path still works next to the device cache
    records = kb.document_top_pairwise_scores(n=2)
    assert (records[0][1]['id'], records[0][2]['id']) == (1, 2)
    kb.close()

    # add / delete invalidate the DEVICE cache too (reference kb.py:1523, 1541)
    kb = svs.KB(path, _embed_words)
    assert texts('... forth ...', 1) == ['first doc']
    with kb.bulk_add_docs() as add_doc:
        assert add_doc('forth doc') == 4
    assert texts('... forth ...', 1) == ['forth doc']
    with kb.bulk_del_docs() as del_doc:
        del_doc(1); del_doc(2); del_doc(4)
    assert texts('... forth ...', 1) == ['second doc']
    assert texts('... forth ...', 10) == ['second doc']          # n > N -> N results
    assert kb.retrieve('... forth ...', 0) == []
    kb.close()


def test_kb_retrieve_async_with_concurrent_retrieves(svs_patched, tmp_path):
    svs = svs_patched
    path = str(tmp_path / "adb.sqlite")
    d = 96

    async def embed(texts):
        return [stub_vector(t, d) for t in texts]

    async def go():
        kb = svs.AsyncKB(path, embed)
        async with kb.bulk_add_docs() as add_doc:
            for i in range(500):
                await add_doc(f"doc {i}")
        await kb.load()                                           # pre-warm (kb.py:964-967) -> device
        assert kb.embeddings_matrix.device._matrix is not None
        # many retrieves in flight at once: compute runs in the default executor without the KB lock
        outs = await asyncio.gather(*[kb.retrieve(f"doc {i}", 5) for i in range(24)])
        for i, res in enumerate(outs):
            assert res[0]['doc']['text'] == f"doc {i}" and res[0]['score'] == pytest.approx(1.0, abs=1e-5)
        # a writer invalidates while readers are in flight; everyone still gets a consistent answer
        async def writer():
            async with kb.bulk_add_docs() as add_doc:
                await add_doc("late doc")
        res = await asyncio.gather(writer(), *[kb.retrieve(f"doc {i}", 3) for i in range(8)])
        for i, r in enumerate(res[1:]):
            assert r[0]['doc']['text'] == f"doc {i}"
        assert (await kb.retrieve("late doc", 1))[0]['doc']['text'] == "late doc"
        await kb.close()
    asyncio.run(go())


def test_dropin_matches_the_unpatched_reference_on_a_real_kb(svs_patched, tmp_path):
    """Same SQLite file, same queries: patched KB (GPU) vs the reference's own NumPy path (oracle)."""
    svs = svs_patched
    import sqlite3
    path = str(tmp_path / "kb.sqlite")
    d = 256

    async def embed(texts):
        return [stub_vector(t, d) for t in texts]
    kb = svs.KB(path, embed)
    with kb.bulk_add_docs() as add_doc:
        for i in range(3000):
            add_doc(f"joke number {i}", no_embedding=(i % 50 == 7))
    with kb.bulk_del_docs() as del_doc:
        for i in (5, 6, 100, 2999):
            del_doc(i)
    conn = sqlite3.connect(path)
    m, ids = oracle.build_embeddings_matrix(conn)
    emb_to_doc = dict(conn.execute("SELECT embedding, id FROM docs WHERE embedding IS NOT NULL;").fetchall())
    for qtext, n in [("joke number 17", 10), ("anything else", 100), ("joke", 1000), ("zzz", len(ids))]:
        res = kb.retrieve(qtext, n)
        q = oracle.query_vec_of(stub_vector(qtext, d))
        want = oracle.superheavy(m, ids, q, n)
        doc_to_emb = {v: k for k, v in emb_to_doc.items()}
        got = [(r['score'], doc_to_emb[r['doc']['id']]) for r in res]
        oracle.compare_retrieval(got, want, oracle.scores_of(m, q), ids)
    kb.close()


def test_retrieve_many_equals_looping_retrieve(svs_patched, tmp_path):
    """The additive batched API: result j of retrieve_many == retrieve(queries[j]) (sync and async), on a KB large
    enough to take the tensor-core path (>= 4096 rows)."""
    svs = svs_patched
    path = str(tmp_path / "many.sqlite")
    d = 128

    async def embed(texts):
        return [stub_vector(t, d) for t in texts]
    kb = svs.KB(path, embed)
    with kb.bulk_add_docs() as add_doc:
        for i in range(5000):
            add_doc(f"passage {i}")
    queries = [f"passage {i * 37}" for i in range(20)] + ["something unseen", "another one"]
    many = kb.retrieve_many(queries, 7)
    assert len(many) == len(queries)
    for qtext, res in zip(queries, many):
        single = kb.retrieve(qtext, 7)
        assert [(r['score'], r['doc']['id']) for r in res] == [(r['score'], r['doc']['id']) for r in single]
    assert many[0][0]['doc']['text'] == "passage 0"
    assert kb.retrieve_many([], 5) == []
    assert kb.retrieve_many(queries[:3], 0) == [[], [], []]
    kb.close()

    async def go():
        akb = svs.AsyncKB(path, embed)
        res = await akb.retrieve_many(queries, 3)
        for qtext, r in zip(queries, res):
            s = await akb.retrieve(qtext, 3)
            assert [(x['score'], x['doc']['id']) for x in r] == [(x['score'], x['doc']['id']) for x in s]
        await akb.close()
    asyncio.run(go())


def test_pairwise_with_thousands_of_duplicates_is_handed_to_the_reference_path(svs_patched, tmp_path, caplog):
    """Duplicate detection is the typical use of document_top_pairwise_scores: ~3000 identical documents are 4.5M pairs
    at score 1.0, more than the engine's candidate list holds (SVSB_E_NOMEM).  The reference's np.dot(M, M.T) +
    get_top_pairs answers it, so the patched call must too (it delegates), not raise."""
    import logging
    svs = svs_patched
    d = 32

    async def embed(texts):
        return [stub_vector(t.split("#")[0], d) for t in texts]
    kb = svs.KB(str(tmp_path / "dups.sqlite"), embed)
    with kb.bulk_add_docs() as add_doc:
        for i in range(3100):
            add_doc(f"the same text#{i}")
        for i in range(40):
            add_doc(f"other {i}")
    with caplog.at_level(logging.INFO):
        res = kb.document_top_pairwise_scores(n=25)
    assert len(res) == 25
    for score, a, b in res:
        assert score == pytest.approx(1.0, abs=1e-5) and a['text'].startswith("the same text") and b['text'].startswith("the same text")
    assert any("engine declined the pairwise query" in r.getMessage() for r in caplog.records)
    # a query the engine does answer still goes through it
    with kb.bulk_del_docs() as del_doc:
        for i in range(3, 3101):
            del_doc(i)
    res = kb.document_top_pairwise_scores(n=3)
    assert res[0][0] == pytest.approx(1.0, abs=1e-5) and {res[0][1]['id'], res[0][2]['id']} == {1, 2}
    kb.close()
