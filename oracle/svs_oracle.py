"""CPU oracle for the SVS retrieve hot path.  TEST INFRASTRUCTURE ONLY.

This module is a plain NumPy restatement of what the reference (Rhobota/svs 0.7.4)
computes on the path behind ``KB.retrieve`` / ``AsyncKB.retrieve``.  It is *not* part of
the product: only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it.  ``svs_b200`` never does.

Where the arithmetic really lives: the reference's hot path is two calls into a
third-party dependency that is not under ``/root/reference`` -- NumPy (``pyproject.toml:39``,
unpinned; this image resolves it to NumPy 2.3.x + OpenBLAS 0.3.30).  ``np.dot`` on a
C-contiguous float32 (N, D) matrix and a float32 (D,) vector is a BLAS ``sgemv`` with a float32
result; ``np.argpartition`` is an introselect.  The restatement below therefore *calls the same
NumPy entry points* the reference calls, in the same order, so on a given machine the oracle's
scores are bit-identical to the reference's.

Parity pinning (see ``oracle/make_golden.py`` and ``tests/test_oracle.py``):
  * every known-answer case of the reference's own ``tests/test_util.py:142-400``
    (``get_top_k``), ``tests/test_embeddings.py:13-22`` (blob codec) and
    ``tests/test_kb.py:753-808`` (matrix build) is restated as a test;
  * golden vectors under ``tests/golden/`` were produced by importing the *real* reference from
    ``/root/reference/src`` in the build container (script committed: ``oracle/make_golden.py``).

Each function cites the reference file:line it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import sqlite3
import struct
from typing import List, Sequence, Tuple

import numpy as np

# src/svs/kb.py:58 -- tolerance of the unit-norm guard applied to docs and queries.
EMBEDDING_MAGNITUDE_TOLERANCE = 0.001

# BASELINE.json north_star: "<=1e-5 relative score error".  Cosine scores live in [-1, 1] and a score
# near zero is the result of cancellation: BOTH fp32 implementations (OpenBLAS and the CUDA kernel)
# then carry an absolute rounding error of a few ulp(1.0) ~ 1e-7 however small the score is, so a purely
# relative bound is meaningless there.  The comparator therefore accepts |diff| <= rtol*|score| + atol
# with atol = 1e-6, i.e. ten times tighter than 1e-5 relative to the cosine scale of 1.0.
SCORE_RTOL = 1e-5
SCORE_ATOL = 1e-6


# --------------------------------------------------------------------------------------
# Blob codec -- src/svs/embeddings/util.py:15-23
# --------------------------------------------------------------------------------------
def embedding_to_bytes(embedding: Sequence[float]) -> bytes:
    """src/svs/embeddings/util.py:15-16: little-endian float32, no header."""
    return struct.pack(f'<{len(embedding)}f', *embedding)


def embedding_from_bytes(embedding: bytes) -> List[float]:
    """src/svs/embeddings/util.py:19-23: inverse of the above; length must divide by 4."""
    size = struct.calcsize('<f')
    assert (len(embedding) % size) == 0
    n_items = len(embedding) // size
    return list(struct.unpack(f'<{n_items}f', embedding))


def magnitude_ok(vectors: Sequence[Sequence[float]], tolerance: float = EMBEDDING_MAGNITUDE_TOLERANCE) -> bool:
    """src/svs/embeddings/util.py:33-38: norms computed in float32; reject |norm-1| > tol."""
    vectors_np = np.array(vectors, dtype=np.float32)
    mags = np.sqrt((vectors_np * vectors_np).sum(axis=1))
    return not bool((np.abs(mags - 1.0) > tolerance).any())


# --------------------------------------------------------------------------------------
# Matrix build -- src/svs/kb.py:573-618
# --------------------------------------------------------------------------------------
def build_embeddings_matrix(conn: sqlite3.Connection) -> Tuple[np.ndarray, np.ndarray]:
    """src/svs/kb.py:573-618 (`_Querier.build_embeddings_matrix`).

    COUNT(*) -> n; first row's blob length -> m (0 when the table is empty); then one
    ``SELECT id, embedding FROM embeddings`` scan (no ORDER BY: rowid order) filling a float32
    (n, m) matrix and an int64 (n,) id vector.  The reference goes blob -> struct.unpack ->
    list -> row assign; float32 -> float64 -> float32 is lossless, so ``np.frombuffer`` gives the
    identical bits (checked against the real reference in tests/golden).
    """
    n = conn.execute("SELECT COUNT(*) FROM embeddings;").fetchone()[0]
    row = conn.execute("SELECT embedding FROM embeddings LIMIT 1;").fetchone()
    m = len(row[0]) // 4 if row is not None else 0
    embeddings_matrix = np.zeros((n, m), dtype=np.float32)
    emb_id_lookup = np.zeros(n, dtype=np.int64)
    i = -1
    for i, (emb_id, blob) in enumerate(conn.execute("SELECT id, embedding FROM embeddings;")):
        vec = np.frombuffer(blob, dtype='<f4')
        assert len(vec) == m           # kb.py:613
        embeddings_matrix[i] = vec
        emb_id_lookup[i] = emb_id
    assert i == n - 1                  # kb.py:616
    return embeddings_matrix, emb_id_lookup


# --------------------------------------------------------------------------------------
# Top-k -- src/svs/util.py:190-203
# --------------------------------------------------------------------------------------
def get_top_k(scores: np.ndarray, top_k: int) -> List[Tuple[float, int]]:
    """src/svs/util.py:190-203.

    k clipped to len (198-199); k<=0 -> [] (200-201); ``np.argpartition(scores, -k)[-k:]``
    (202); then ``sorted(..., reverse=True)`` over (float(score), int(index)) tuples (203), i.e.
    score descending and, among equal scores, index DEscending.
    """
    assert scores.ndim == 1
    assert isinstance(top_k, int)
    if top_k > len(scores):
        top_k = len(scores)
    if top_k <= 0:
        return []
    indices = np.argpartition(scores, -top_k)[-top_k:]
    return sorted([(float(scores[i]), int(i)) for i in indices], reverse=True)


# --------------------------------------------------------------------------------------
# The hot closure -- src/svs/kb.py:1184-1189 (async) and 1622-1627 (sync)
# --------------------------------------------------------------------------------------
def scores_of(embeddings_matrix: np.ndarray, query_vec: np.ndarray) -> np.ndarray:
    """src/svs/kb.py:1185 / 1623: ``x = np.dot(embeddings_matrix, query_vec)`` (float32 out).

    Raises ValueError for a D mismatch or for the empty (0, 0) matrix, exactly as NumPy does
    for the reference (observable edge behaviour).
    """
    return np.dot(embeddings_matrix, query_vec)


def superheavy(embeddings_matrix: np.ndarray, emb_id_lookup: np.ndarray,
               query_vec: np.ndarray, n: int) -> List[Tuple[float, int]]:
    """src/svs/kb.py:1622-1627: dot, top-k, row index -> embeddings.id."""
    x = scores_of(embeddings_matrix, query_vec)
    emb_ids = []
    for score, index in get_top_k(x, n):
        emb_ids.append((score, int(emb_id_lookup[index])))
    return emb_ids


def get_top_pairs(pairwise_scores_as_matrix: np.ndarray, top_k: int) -> List[Tuple[float, int, int]]:
    """src/svs/util.py:206-233: upper triangle without the diagonal (np.triu_indices_from(k=1)) -> get_top_k on the
    flattened values -> (score, row, col).  Exact ties come out by DESCENDING flat index (get_top_k's sort)."""
    assert len(pairwise_scores_as_matrix.shape) == 2
    rows, cols = pairwise_scores_as_matrix.shape
    assert rows == cols
    indices = np.triu_indices_from(pairwise_scores_as_matrix, k=1)
    vals = pairwise_scores_as_matrix[indices]
    top = get_top_k(vals, top_k=top_k)
    return [(score, int(indices[0][ii]), int(indices[1][ii])) for score, ii in top]


def top_pairwise(embeddings_matrix: np.ndarray, emb_id_lookup: np.ndarray, n: int) -> List[Tuple[float, int, int]]:
    """The superheavy() closure of document_top_pairwise_scores, src/svs/kb.py:1650-1656 (sync), 1218-1224 (async)."""
    pairwise = np.dot(embeddings_matrix, embeddings_matrix.T)
    return [(score, int(emb_id_lookup[i1]), int(emb_id_lookup[i2])) for score, i1, i2 in get_top_pairs(pairwise, n)]


def compare_pairs(engine: Sequence[Tuple[float, int, int]], oracle_list: Sequence[Tuple[float, int, int]],
                  pairwise: np.ndarray, emb_id_lookup: np.ndarray,
                  rtol: float = SCORE_RTOL, atol: float = SCORE_ATOL) -> dict:
    """Tolerance-aware comparison of two top-pairs lists (same rules as compare_retrieval): same length; every
    engine score within tolerance of the oracle matrix's entry for that pair; every rank either the same pair or a
    near-tie swap; engine sorted by (score desc, first row asc, second row asc); pairs distinct and i < j in row order."""
    assert len(engine) == len(oracle_list), (len(engine), len(oracle_list))
    row_of = {int(e): r for r, e in enumerate(emb_id_lookup)}
    seen = set()
    exact = 0
    max_rel = 0.0
    prev = None
    for r, ((s, a, b), (so, ao, bo)) in enumerate(zip(engine, oracle_list)):
        i, j = row_of[a], row_of[b]
        assert i < j, f"rank {r}: pair not in the upper triangle: rows {i}, {j}"
        assert (i, j) not in seen, f"rank {r}: duplicate pair"
        seen.add((i, j))
        ref = float(pairwise[i, j])
        tol = rtol * abs(ref) + atol
        assert abs(s - ref) <= tol, f"rank {r}: score {s} vs oracle {ref}"
        max_rel = max(max_rel, abs(s - ref) / max(abs(ref), 1e-30))
        if (a, b) == (ao, bo):
            exact += 1
        else:
            assert abs(ref - so) <= rtol * abs(so) + atol, f"rank {r}: pair ({a},{b}) is not a near-tie of the oracle's ({ao},{bo})"
        key = (-s, i, j)
        assert prev is None or prev <= key, f"rank {r}: not sorted by (score desc, row asc, row asc)"
        prev = key
    return {"n": len(engine), "exact_rank_matches": exact, "max_rel_score_err": max_rel}


def query_vec_of(list_of_floats: Sequence[float]) -> np.ndarray:
    """src/svs/kb.py:1182 / 1620: the provider's float list -> float32, NOT re-normalised."""
    return np.array(list_of_floats, dtype=np.float32)


# --------------------------------------------------------------------------------------
# Canonical order of the new engine (north star: score desc, ties by ascending doc id)
# --------------------------------------------------------------------------------------
def canonical_top_k(scores: np.ndarray, ids: np.ndarray, k: int) -> List[Tuple[float, int]]:
    """Exact top-k of `scores` under the total order (score desc, id asc).

    This is the order the north star specifies for the engine; it
    differs from get_top_k only in how *exactly equal* float32 scores are arranged and in which of
    several boundary-tied rows is kept (the reference leaves that to introselect).  Used to test
    the selection kernel bit-exactly on a given score vector.
    """
    n = len(scores)
    k = min(k, n)
    if k <= 0:
        return []
    order = np.lexsort((ids, -scores.astype(np.float64)))  # last key is primary
    top = order[:k]
    return [(float(scores[i]), int(ids[i])) for i in top]


# --------------------------------------------------------------------------------------
# Tolerance-aware comparator -- SURVEY.md section 8(c)
# --------------------------------------------------------------------------------------
def compare_retrieval(engine: Sequence[Tuple[float, int]],
                      oracle: Sequence[Tuple[float, int]],
                      oracle_scores: np.ndarray,
                      emb_id_lookup: np.ndarray,
                      rtol: float = SCORE_RTOL, atol: float = SCORE_ATOL) -> dict:
    """Check an engine result list against the oracle's on the same (M, q, n).

    (1) same length; (2) every engine score within rtol*|score| + atol of the oracle's score for that
    id; (3) at every rank either the same id, or a near-tie swap: the oracle's score of the engine's
    id is within the same tolerance of the oracle's score at that rank; (4) engine list ordered by
    (score desc, id asc); (5) no duplicate ids.  Returns a dict of diagnostics and raises
    AssertionError with a readable message on the first violation.
    """
    row_of = {int(e): i for i, e in enumerate(emb_id_lookup)}
    assert len(engine) == len(oracle), f"length {len(engine)} != oracle {len(oracle)}"
    exact = 0
    max_rel = 0.0
    seen = set()
    for r, ((es, eid), (os_, oid)) in enumerate(zip(engine, oracle)):
        assert eid in row_of, f"rank {r}: id {eid} is not in the matrix"
        assert eid not in seen, f"rank {r}: duplicate id {eid}"
        seen.add(eid)
        xs = float(oracle_scores[row_of[eid]])
        err = abs(es - xs)
        if abs(xs) > atol / rtol:
            max_rel = max(max_rel, err / abs(xs))
        assert err <= rtol * abs(xs) + atol, f"rank {r}: id {eid} score {es!r} vs oracle {xs!r} (abs err {err:.3e})"
        if eid == oid:
            exact += 1
        else:
            gap = abs(xs - os_)
            assert gap <= rtol * abs(os_) + atol, (
                f"rank {r}: engine id {eid} (oracle score {xs!r}) vs oracle id {oid} "
                f"(score {os_!r}): not a near-tie (gap {gap:.3e})")
        if r > 0:
            ps, pid = engine[r - 1]
            assert (ps > es) or (ps == es and pid < eid), (
                f"rank {r}: order violated: ({ps!r},{pid}) then ({es!r},{eid})")
    return {"n": len(engine), "exact_rank_matches": exact, "max_rel_score_err": max_rel}


# --------------------------------------------------------------------------------------
# Synthetic inputs of the BASELINE.json configs -- SURVEY.md section 8(d)
# --------------------------------------------------------------------------------------
def synth_matrix_uniform(n: int, d: int, seed: int = 0) -> np.ndarray:
    """examples/One Million Documents Benchmark.ipynb:66-71: uniform [0,1) rows / L2 norm, in f32."""
    rng = np.random.default_rng(seed)
    m = rng.random((n, d), dtype=np.float32)
    m /= np.sqrt((m.astype(np.float64) ** 2).sum(axis=1)).astype(np.float32)[:, None]
    return m


def synth_matrix_normal(n: int, d: int, seed: int = 0) -> np.ndarray:
    """Second distribution of SURVEY 8(d): standard-normal rows / L2 norm (scores ~ N(0, 1/sqrt(D)))."""
    rng = np.random.default_rng(seed)
    m = rng.standard_normal((n, d), dtype=np.float32)
    m /= np.sqrt((m.astype(np.float64) ** 2).sum(axis=1)).astype(np.float32)[:, None]
    return m


def synth_queries(nq: int, d: int, seed: int = 1, dist: str = "uniform") -> np.ndarray:
    rng = np.random.default_rng(seed)
    if dist == "uniform":
        q = rng.random((nq, d), dtype=np.float32)
    else:
        q = rng.standard_normal((nq, d), dtype=np.float32)
    q /= np.sqrt((q.astype(np.float64) ** 2).sum(axis=1)).astype(np.float32)[:, None]
    return q


# Counter-based generator shared bit-for-bit with the CUDA side (svs_b200/csrc/synth.cu):
# element (row, col) of the matrix seeded `seed` is a pure function of (seed, row, col), so any
# slab of the 10M-row config can be regenerated on the host without a 61 GB copy.
_M1 = np.uint64(0xFF51AFD7ED558CCD)
_M2 = np.uint64(0xC4CEB9FE1A85EC53)
_GOLD = np.uint64(0x9E3779B97F4A7C15)


def _mix64(z: np.ndarray) -> np.ndarray:
    with np.errstate(over='ignore'):
        z = z ^ (z >> np.uint64(33))
        z = z * _M1
        z = z ^ (z >> np.uint64(33))
        z = z * _M2
        z = z ^ (z >> np.uint64(33))
    return z


def counter_uniform_rows(seed: int, row0: int, nrows: int, d: int) -> np.ndarray:
    """Un-normalised uniform [0,1) float32 values: u = (hash(seed,row,col) >> 40) * 2**-24."""
    rows = np.arange(row0, row0 + nrows, dtype=np.uint64)[:, None]
    cols = np.arange(d, dtype=np.uint64)[None, :]
    with np.errstate(over='ignore'):
        ctr = (rows * np.uint64(d) + cols) * _GOLD + np.uint64(seed) * _M2 + np.uint64(1)
    h = _mix64(ctr)
    return ((h >> np.uint64(40)).astype(np.float32)) * np.float32(2.0 ** -24)
