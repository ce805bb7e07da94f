"""Generate tests/golden/* by running the REAL reference.  TEST INFRASTRUCTURE ONLY.

Run in the build container (where /root/reference is mounted):

    python oracle/make_golden.py

It imports ``svs`` from ``/root/reference/src`` (never from this repo), drives the reference's
own functions on seeded inputs and stores inputs *and* outputs, so that the tests can pin
``oracle/svs_oracle.py`` (CPU suite) and the CUDA engine (GPU suite) against the reference
without the reference being present.  Fixtures are kept small (a few MB in total).

Fixtures written:
  topk_cases.npz / topk_cases.json   -- svs.util.get_top_k (src/svs/util.py:190-203) on the
                                        reference's own test inputs (tests/test_util.py:142-400)
                                        and on seeded float32 vectors;
  superheavy_d96.npz, superheavy_d1536.npz
                                     -- np.dot + get_top_k + emb_id_lookup (src/svs/kb.py:1622-1627);
  codec.json                         -- embedding_to_bytes / from_bytes (src/svs/embeddings/util.py:15-23);
  kb_small.sqlite + kb_small.json + kb_small_matrix.npz
                                     -- a real SQLite KB written by svs.KB.bulk_add_docs /
                                        bulk_del_docs, the matrix the reference builds from it
                                        (src/svs/kb.py:573-618) and what svs.KB.retrieve returns.
"""
from __future__ import annotations

import itertools
import json
import os
import sys
import zlib

import numpy as np

REF_SRC = "/root/reference/src"
sys.path.insert(0, REF_SRC)
import svs                                            # noqa: E402  (the real reference)
from svs.util import get_top_k                        # noqa: E402
from svs.embeddings.util import embedding_to_bytes, embedding_from_bytes  # noqa: E402
from svs.kb import _DB                                # noqa: E402

assert svs.__file__.startswith(REF_SRC), svs.__file__

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
OUT = os.path.normpath(OUT)
os.makedirs(OUT, exist_ok=True)


def stub_vector(text: str, d: int) -> list:
    """Deterministic near-unit vector for a text (same helper lives in tests/_stubs.py)."""
    rng = np.random.default_rng(zlib.crc32(text.encode("utf-8")))
    v = rng.standard_normal(d)
    v /= np.sqrt((v * v).sum())
    return [float(x) for x in v]


def gen_topk():
    arrays = {}
    cases = []

    def add(name, scores, k):
        arrays[name] = scores
        cases.append({"scores": name, "k": k,
                      "expected": [[s, i] for s, i in get_top_k(scores, k)]})

    # the reference's own known-answer inputs: every permutation of up to 3 distinct scores
    base = [0.4, 0.2, 0.9]
    idx = 0
    for n in range(0, 4):
        for perm in sorted(set(itertools.permutations(base[:n]))):
            a = np.array(perm, dtype=np.float64)
            name = f"perm{idx}"
            idx += 1
            for k in range(0, 5):
                add(name, a, k)
    # seeded float32 vectors, distinct values (no ties -> result independent of introselect)
    for j, (n, k) in enumerate([(10, 3), (100, 10), (1000, 100), (5000, 1000), (4096, 4096),
                                (777, 1), (20000, 100), (20000, 1000), (65, 64), (3, 7)]):
        rng = np.random.default_rng(100 + j)
        a = np.unique(rng.standard_normal(n).astype(np.float32))   # distinct values
        rng.shuffle(a)
        add(f"rand{j}", a, k)
    # the hard, clustered distribution of the 1M benchmark (scores ~ 0.752 +- 0.007)
    rng = np.random.default_rng(7)
    a = (0.752 + 0.0072 * rng.standard_normal(30000)).astype(np.float32)
    a = np.unique(a)
    rng.shuffle(a)
    add("clustered", a, 100)
    add("clustered", a, 1000)
    # negative / mixed sign / special magnitudes
    a = np.array([-1.0, -0.0, 0.0, 1e-38, -1e-38, 3.0, -3.0, 1e30, -1e30], dtype=np.float32)
    add("signs", a, 4)
    add("signs", a, 9)
    np.savez_compressed(os.path.join(OUT, "topk_cases.npz"), **arrays)
    with open(os.path.join(OUT, "topk_cases.json"), "w") as f:
        json.dump(cases, f)
    print(f"topk: {len(cases)} cases, {len(arrays)} arrays")


def gen_superheavy():
    for tag, n, d, nq, ks in [("d96", 3000, 96, 8, [1, 10, 100, 1000, 3000, 5000]),
                              ("d1536", 300, 1536, 4, [10, 100])]:
        rng = np.random.default_rng(2024)
        m = rng.random((n, d), dtype=np.float32)
        m /= np.sqrt((m * m).sum(axis=1))[:, None]
        # embeddings.id with gaps, ascending (as a rowid scan yields)
        ids = np.cumsum(rng.integers(1, 4, size=n)).astype(np.int64)
        q = rng.standard_normal((nq, d)).astype(np.float32)
        q /= np.sqrt((q * q).sum(axis=1))[:, None]
        out = {"matrix": m, "emb_ids": ids, "queries": q, "ks": np.array(ks)}
        for qi in range(nq):
            x = np.dot(m, q[qi])                          # kb.py:1623
            assert x.dtype == np.float32
            out[f"scores_q{qi}"] = x
            for k in ks:
                res = [(s, int(ids[i])) for s, i in get_top_k(x, k)]   # kb.py:1625-1626
                out[f"top_q{qi}_k{k}_scores"] = np.array([s for s, _ in res], dtype=np.float64)
                out[f"top_q{qi}_k{k}_ids"] = np.array([e for _, e in res], dtype=np.int64)
        np.savez_compressed(os.path.join(OUT, f"superheavy_{tag}.npz"), **out)
        print(f"superheavy_{tag}: n={n} d={d}")


def gen_codec():
    cases = []
    rng = np.random.default_rng(5)
    for vec in ([], [1.0], [1.0, 3.5], [float(x) for x in rng.standard_normal(7)]):
        b = embedding_to_bytes(vec)
        cases.append({"vector": vec, "hex": b.hex(), "roundtrip": embedding_from_bytes(b)})
    with open(os.path.join(OUT, "codec.json"), "w") as f:
        json.dump(cases, f)
    print(f"codec: {len(cases)} cases")


def gen_kb():
    d = 64
    path = os.path.join(OUT, "kb_small.sqlite")
    for p in (path,):
        if os.path.exists(p):
            os.remove(p)

    async def embedding_func(texts):
        return [stub_vector(t, d) for t in texts]

    kb = svs.KB(path, embedding_func)
    texts = [f"document number {i}: {'lorem ipsum ' * (i % 5)}#{i * 7919 % 1000}" for i in range(420)]
    with kb.bulk_add_docs() as add_doc:
        for i, t in enumerate(texts):
            if i % 97 == 5:
                add_doc(t, no_embedding=True)       # kb.py:1507-1508: no row in `embeddings`
            else:
                add_doc(t, meta={"i": i} if i % 3 == 0 else None)
    with kb.bulk_del_docs() as del_doc:              # leaves gaps in embeddings.id
        for doc_id in (3, 50, 51, 52, 200, 419):
            del_doc(doc_id)
    with kb.bulk_add_docs() as add_doc:              # new ids continue after the max
        for i in range(5):
            add_doc(f"late addition {i}")
    queries = ["what is document number 17", "lorem ipsum", "late addition", "zzz", "#123"]
    expected = {"d": d, "queries": []}
    for qtext in queries:
        for n in (1, 10, 100, 1000):
            res = kb.retrieve(qtext, n)
            expected["queries"].append({
                "text": qtext, "vector": stub_vector(qtext, d), "n": n,
                "results": [{"score": r["score"], "doc_id": r["doc"]["id"], "text": r["doc"]["text"]}
                            for r in res]})
    expected["len"] = len(kb)
    kb.close()
    # the matrix the reference itself builds from that file
    db = _DB(path)
    with db as q:
        m, ids = q.build_embeddings_matrix()
    db.close()
    np.savez_compressed(os.path.join(OUT, "kb_small_matrix.npz"), matrix=m, emb_ids=ids)
    with open(os.path.join(OUT, "kb_small.json"), "w") as f:
        json.dump(expected, f)
    print(f"kb_small: {m.shape} ids {ids[0]}..{ids[-1]}, {len(expected['queries'])} retrieves")


if __name__ == "__main__":
    gen_topk()
    gen_superheavy()
    gen_codec()
    gen_kb()
    total = sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT))
    print(f"golden dir: {total / 1e6:.2f} MB")
