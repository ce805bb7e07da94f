"""Load-path measurement: SQLite `embeddings` scan -> engine.

    python scripts/load_bench.py [rows] [dims] [--fake]

Builds a scratch SQLite file with the reference's `embeddings` table (src/svs/kb.py:80-83, blobs as
src/svs/embeddings/util.py:15-16 packs them), then times
  * the reference's own per-row decode (src/svs/kb.py:603-616) on a bounded sample of rows,
  * svs_b200.load_from_connection (rows -> pinned slabs -> device), the whole table,
and checks the device matrix against the blobs.  --fake: no GPU, slabs are plain host arrays (host side only).
"""
import os
import sqlite3
import struct
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

args = [a for a in sys.argv[1:] if not a.startswith("--")]
n = int(args[0]) if len(args) > 0 else 100_000
d = int(args[1]) if len(args) > 1 else 1536
fake = "--fake" in sys.argv

tmp = tempfile.mkdtemp(prefix="svsb_load_")
path = os.path.join(tmp, "kb.sqlite")
conn = sqlite3.connect(path, isolation_level=None, check_same_thread=False)
conn.execute("CREATE TABLE embeddings (id INTEGER PRIMARY KEY, embedding BLOB NOT NULL) STRICT;")
# the part of the reference's docs table the loader's count uses (kb.py:85-96)
conn.execute("CREATE TABLE docs (id INTEGER PRIMARY KEY, embedding INTEGER REFERENCES embeddings(id));")
conn.execute("CREATE INDEX idx_docs_embedding ON docs(embedding);")
rng = np.random.default_rng(0)
t0 = time.perf_counter()
conn.execute("BEGIN;")
step = 4096
for a in range(0, n, step):
    m = rng.random((min(step, n - a), d), dtype=np.float32)
    m /= np.sqrt((m * m).sum(axis=1))[:, None]
    conn.executemany("INSERT INTO embeddings (id, embedding) VALUES (?, ?);", ((a + i + 1, m[i].tobytes()) for i in range(len(m))))
    conn.executemany("INSERT INTO docs (embedding) VALUES (?);", ((a + i + 1,) for i in range(len(m))))
conn.execute("COMMIT;")
print(f"built {n} x {d} ({n * d * 4 / 1e9:.2f} GB of blobs, file {os.path.getsize(path) / 1e9:.2f} GB) in {time.perf_counter() - t0:.1f}s")

# the reference's decode on a bounded sample (kb.py:603-616 + embeddings/util.py:19-23)
sample = min(n, 5000)
t0 = time.perf_counter()
mat = np.zeros((sample, d), dtype=np.float32)
for i, (emb_id, blob) in enumerate(conn.execute(f"SELECT id, embedding FROM embeddings LIMIT {sample};")):
    mat[i] = list(struct.unpack(f"<{len(blob) // 4}f", blob))
ref_s = time.perf_counter() - t0
print(f"reference decode: {sample} rows in {ref_s:.2f}s = {sample / ref_s:.0f} rows/s = {sample * d * 4 / ref_s / 1e9:.3f} GB/s"
      f"  (-> {n / (sample / ref_s):.0f}s for {n} rows)")

import svs_b200
from svs_b200 import matrix as matrix_mod
if fake:
    class Fake:
        def load_begin(self, n, d, normalize=False):
            self.n, self.d, self.f = n, d, 0
            self.slab = np.zeros(5461 * d * 4, np.uint8); self.ids = np.zeros(5461, np.int64)
        def acquire_slab(self, d):
            cap = min(5461, self.n - self.f); return self.slab[:cap * d * 4], self.ids[:cap]
        def commit_slab(self, c): self.f += c
        def load_end(self): return 1
        def load_abort(self): pass
        def snapshot(self):
            class S: shape = (0, 0); generation = 1
            return S()
    eng = Fake()
else:
    eng = svs_b200.Engine()
for rep in range(2):
    t0 = time.perf_counter()
    dm = matrix_mod.load_from_connection(eng, conn)
    dt = time.perf_counter() - t0
    print(f"svs_b200 load_from_connection (scan through the Python connection): {n} rows in {dt:.2f}s = {n / dt:.0f} rows/s = "
          f"{n * d * 4 / dt / 1e9:.3f} GB/s  ({(sample / ref_s) and (n / dt) / (sample / ref_s):.1f}x the reference decode)")
if not fake:
    # the native scan (svsb_load_sqlite: libsqlite3 bound at run time, T read-only connections over rowid ranges)
    for threads in (1, 2, 4, 8, 16, 0):
        t0 = time.perf_counter()
        eng.load_sqlite(path, threads=threads)
        dt = time.perf_counter() - t0
        print(f"svs_b200 svsb_load_sqlite threads={threads or 'default'}: {n} rows in {dt:.2f}s = {n / dt:.0f} rows/s = "
              f"{n * d * 4 / dt / 1e9:.3f} GB/s", flush=True)
    from svs_b200.engine import sqlite_read
    t0 = time.perf_counter()
    hm, hid = sqlite_read(path, 0)
    dt = time.perf_counter() - t0
    print(f"svsb_sqlite_read (host arrays only, default threads): {n * d * 4 / dt / 1e9:.3f} GB/s")
if not fake:
    rows, ids = eng.read_rows(0, min(n, 5000))
    assert rows.tobytes() == mat[:len(rows)].tobytes() and (ids == np.arange(1, len(ids) + 1)).all(), "device matrix != blobs"
    print("device matrix equals the blobs bit for bit; norm stats:", eng.norm_stats())
    eng.close()
conn.close()
os.remove(path); os.rmdir(tmp)
