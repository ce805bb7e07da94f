"""BASELINE.json configs[0] end to end through the drop-in boundary: a real SQLite KB of 10,548 documents x 1536-d
(the dad-jokes shape; SURVEY.md section 8d recipe: stub embedding func, document i -> default_rng([1, i]) unit vector,
any other text -> the fixed query vector from default_rng([1, 2**31])) built through the reference's own
KB.bulk_add_docs, then KB.retrieve(query, n=10) timed with the UNMODIFIED reference (NumPy path) and with
svs_b200.install() applied.  Whole-call times: embedding stub + similarity + top-n + the SQL fetch of the n documents.

    python scripts/c1_dropin_bench.py [docs] [dims] [n]        # needs oracle/_ref (python oracle/build_ref.py)
"""
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref", "svs_ref.bin"))     # the byte-compiled reference

docs = int(sys.argv[1]) if len(sys.argv) > 1 else 10_548
d = int(sys.argv[2]) if len(sys.argv) > 2 else 1536
n = int(sys.argv[3]) if len(sys.argv) > 3 else 10
texts = [f"dad joke number {i}: why did the vector cross the hyperplane?" for i in range(docs)]
index_of = {t: i for i, t in enumerate(texts)}


def unit(seed_pair):
    v = np.random.default_rng(seed_pair).standard_normal(d)
    return (v / np.sqrt((v * v).sum())).tolist()


async def embed(batch):
    return [unit([1, index_of[t]]) if t in index_of else unit([1, 2 ** 31]) for t in batch]


def timed(kb, label, reps):
    t0 = time.perf_counter()
    first = kb.retrieve("what is funny?", n=n)
    t_first = time.perf_counter() - t0
    lat = []
    for _ in range(reps):
        t0 = time.perf_counter()
        res = kb.retrieve("what is funny?", n=n)
        lat.append(time.perf_counter() - t0)
    assert [r["doc"]["id"] for r in res] == [r["doc"]["id"] for r in first]
    lat = np.array(lat) * 1e3
    return {"impl": label, "first_query_s": round(t_first, 3), "warm_ms_median": round(float(np.median(lat)), 4),
            "warm_ms_min": round(float(lat.min()), 4), "calls": reps,
            "top": [(round(r["score"], 6), r["doc"]["id"]) for r in res[:3]]}, res


def main():
    import svs
    path = os.path.join(tempfile.mkdtemp(), "dad_jokes_shape.sqlite")
    t0 = time.perf_counter()
    kb = svs.KB(path, embed)
    with kb.bulk_add_docs() as add_doc:
        for t in texts:
            add_doc(t)
    kb.close()
    build_s = time.perf_counter() - t0
    out = {"workload": f"{docs} docs x {d}-d fp32 in SQLite, KB.retrieve(query, n={n}), stub embedder (no network)",
           "kb_build_s": round(build_s, 2), "sqlite_mb": round(os.path.getsize(path) / 1e6, 1), "cores": os.cpu_count()}
    kb = svs.KB(path, embed)
    ref, ref_res = timed(kb, "reference (NumPy path, unmodified)", 30)
    kb.close()
    import svs_b200
    svs_b200.install(svs)
    kb = svs.KB(path, embed)
    ours, our_res = timed(kb, "svs_b200.install(svs): libsvsb200.so on cuda:0", 300)
    kb.close()
    svs_b200.uninstall()
    same_ids = [r["doc"]["id"] for r in ref_res] == [r["doc"]["id"] for r in our_res]
    err = max(abs(a["score"] - b["score"]) / abs(a["score"]) for a, b in zip(ref_res, our_res))
    out.update({"reference": ref, "ours": ours, "same_doc_ids_and_order": same_ids, "max_rel_score_diff": err,
                "warm_speedup": round(ref["warm_ms_median"] / ours["warm_ms_median"], 1),
                "first_query_speedup": round(ref["first_query_s"] / ours["first_query_s"], 1)})
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
