"""Device-resident mirror of the reference's vector cache and matrix build.

`DeviceEmbeddingsMatrix` has the interface of `svs.kb._EmbeddingsMatrix` (reference
src/svs/kb.py:856-893): `get_sync(db)`, `async get(db)`, `invalidate()`.  Instead of two NumPy arrays
it hands back a `DeviceMatrix` handle whose `.retrieve(query_vec, n)` is the reference's `superheavy()`
closure (src/svs/kb.py:1622-1627) executed on the GPU.

`load_from_connection` is `_Querier.build_embeddings_matrix` (src/svs/kb.py:573-618): same three
statements against the same schema (src/svs/kb.py:80-83), same scan order, same asserts -- but the
blobs go straight into the engine's pinned staging slabs and on to the device, never through
`struct.unpack` lists.
"""
from __future__ import annotations

import asyncio
import logging
import sqlite3
import threading
from typing import Any, List, Optional, Sequence, Tuple

import numpy as np

from .engine import Engine, Snapshot

_LOG = logging.getLogger(__name__)


class DeviceMatrix:
    """Handle on one loaded generation: what `embeddings_matrix, emb_id_lookup` are to the reference."""

    def __init__(self, engine: Engine):
        self._engine = engine
        self._snap: Snapshot = engine.snapshot()    # pins this generation until the handle dies
        self.shape = self._snap.shape               # (N, D); (0, 0) for an empty table, like the reference
        self.generation = self._snap.generation

    def __len__(self) -> int:
        return self.shape[0]

    def retrieve(self, query_vec: np.ndarray, n: int) -> List[Tuple[float, int]]:
        """superheavy(): [(score, emb_id)] of the n best rows.  Raises ValueError for a D mismatch or
        an empty matrix exactly where the reference's np.dot does."""
        return self._snap.retrieve(query_vec, n)

    def top_pairs(self, n: int) -> List[Tuple[float, int, int]]:
        """The `superheavy()` of document_top_pairwise_scores (src/svs/kb.py:1650-1656): np.dot(M, M.T) + get_top_pairs
        + emb_id_lookup, on the GPU, without materialising the N x N score matrix."""
        return self._snap.top_pairs(n)

    def retrieve_many(self, query_vecs: np.ndarray, n: int) -> List[List[Tuple[float, int]]]:
        """superheavy() for a batch of query vectors in one engine call (bit-identical to looping `retrieve`);
        large batches run as one tensor-core contraction (include/svsb200.h: svsb_query_batch)."""
        return self._snap.retrieve_many(query_vecs, n)


def load_from_connection(engine: Engine, conn: sqlite3.Connection, normalize: bool = False) -> DeviceMatrix:
    """build_embeddings_matrix (src/svs/kb.py:573-618) into the device cache."""
    n = conn.execute("SELECT COUNT(*) FROM embeddings;").fetchone()[0]
    assert isinstance(n, int)
    first = conn.execute("SELECT embedding FROM embeddings LIMIT 1;").fetchone()
    if first is not None:
        assert len(first[0]) % 4 == 0                       # embedding_from_bytes, embeddings/util.py:20-21
        d = len(first[0]) // 4
    else:
        d = 0
    engine.load_begin(n, d, normalize)
    try:
        loaded = 0
        if n and d:
            row_bytes = d * 4
            cur = conn.execute("SELECT id, embedding FROM embeddings;")
            # One memcpy per row, straight from the blob SQLite hands out into the pinned slab (measured: ~2x the
            # throughput of fetchmany + b"".join, which copies every byte three times).
            slab, slab_ids = engine.acquire_slab(d)
            cap, mv, idl, i = len(slab_ids), memoryview(slab), [], 0
            for emb_id, blob in cur:
                if cap == 0:
                    # rows beyond COUNT(*): the reference's `assert i == n-1` (kb.py:616)
                    raise AssertionError("more embedding rows than COUNT(*) reported")
                # every row must have the first row's length (kb.py:613)
                assert len(blob) == row_bytes, "embedding rows of unequal length"
                mv[i * row_bytes:(i + 1) * row_bytes] = blob
                idl.append(emb_id)
                i += 1
                if i == cap:
                    slab_ids[:i] = idl
                    del mv
                    engine.commit_slab(i)
                    loaded += i
                    slab, slab_ids = engine.acquire_slab(d)
                    cap, mv, idl, i = len(slab_ids), memoryview(slab), [], 0
            if i:
                slab_ids[:i] = idl
            del mv
            engine.commit_slab(i)
            loaded += i
        elif n:
            # zero-length blobs: nothing to copy, but ids still define N
            ids = np.fromiter((r[0] for r in conn.execute("SELECT id FROM embeddings;")), dtype=np.int64)
            engine.load_rows(np.zeros((len(ids), 0), dtype=np.float32), ids)
            loaded = len(ids)
        assert loaded == n, f"{loaded} embedding rows scanned, COUNT(*) said {n}"   # kb.py:616
        engine.load_end()
    except BaseException:
        # abandon the half-built generation; the previous one (if any) stays untouched
        try:
            engine.load_abort()
        except Exception:
            pass
        raise
    return DeviceMatrix(engine)


class DeviceEmbeddingsMatrix:
    """Drop-in for `_EmbeddingsMatrix` (src/svs/kb.py:856-893)."""

    def __init__(self, devices: Optional[Sequence[int]] = None, normalize: bool = False) -> None:
        self._devices = list(devices) if devices is not None else None
        self._normalize = normalize
        self._engine: Optional[Engine] = None        # created lazily on first load
        self._matrix: Optional[DeviceMatrix] = None
        self._mu = threading.Lock()

    def _get_engine(self) -> Engine:
        if self._engine is None:
            self._engine = Engine(self._devices)
        return self._engine

    def prewarm(self) -> threading.Thread:
        """Create the engine (CUDA context, streams) on a daemon thread; a load that arrives first simply waits for it."""
        def work() -> None:
            try:
                with self._mu:
                    self._get_engine()
            except Exception as ex:                      # reported again, loudly, by the first load
                _LOG.warning("engine pre-warm failed: %s", ex)
        t = threading.Thread(target=work, name="svs_b200-prewarm", daemon=True)
        t.start()
        return t

    def invalidate(self) -> None:
        """kb.py:861-864.  In-flight queries keep the generation they started on."""
        _LOG.info("invalidating cached device vectors; they'll be re-built next time you `retrieve()`")
        with self._mu:
            self._matrix = None
            if self._engine is not None:
                self._engine.invalidate()

    def _build(self, q: Any) -> DeviceMatrix:
        conn = q.conn if hasattr(q, "conn") else q
        with self._mu:
            engine = self._get_engine()
        return load_from_connection(engine, conn, self._normalize)

    def get_sync(self, db: Any) -> DeviceMatrix:
        """kb.py:866-877: cached handle, or rebuild inside a DB transaction."""
        m = self._matrix
        if m is not None:
            _LOG.info("using cached device vectors")
            return m
        _LOG.info("re-building cached device vectors...")
        with db as q:
            m = self._build(q)
        _LOG.info("re-building cached device vectors... DONE!")
        self._matrix = m
        return m

    async def get(self, db: Any) -> DeviceMatrix:
        """kb.py:879-893: same, with the build in the default executor."""
        m = self._matrix
        if m is not None:
            _LOG.info("using cached device vectors")
            return m
        _LOG.info("re-building cached device vectors...")

        def heavy() -> DeviceMatrix:
            with db as q:
                return self._build(q)
        loop = asyncio.get_running_loop()
        m = await loop.run_in_executor(None, heavy)
        _LOG.info("re-building cached device vectors... DONE!")
        self._matrix = m
        return m

    def close(self) -> None:
        with self._mu:
            self._matrix = None
            if self._engine is not None:
                self._engine.close()
                self._engine = None
