"""GPU: svsb_apply_mutations (append + tombstone) against a FRESH FULL REBUILD of the same
table state -- which is what the reference does after every bulk add / delete (src/svs/kb.py:1062, 1086, 1523, 1541,
573-618).  Bit-exact: same live rows in the same order (read_rows), same (score, id) lists (retrieve), for every k
path: fused selection, full sort (k > 2048), batches."""
import numpy as np
import pytest

from _util import oracle

pytestmark = pytest.mark.gpu


class Mirror:
    """Host model of the `embeddings` table in scan (rowid) order."""

    def __init__(self, rows, ids):
        self.rows, self.ids = rows.copy(), ids.copy()

    def delete(self, del_ids):
        keep = ~np.isin(self.ids, del_ids)
        self.rows, self.ids = self.rows[keep], self.ids[keep]

    def append(self, rows, ids):
        self.rows = np.concatenate([self.rows, rows]); self.ids = np.concatenate([self.ids, ids])

    def next_id(self):
        return int(self.ids.max()) + 1 if len(self.ids) else 1     # SQLite: max(rowid) + 1


def _same(a, b):
    return len(a[0]) == len(b[0]) and np.array_equal(a[0].view(np.uint32), b[0].view(np.uint32)) and np.array_equal(a[1], b[1])


def _compare(inc, fresh, mirror, qs, ks, batch=False):
    fresh.load(mirror.rows, mirror.ids)
    n = len(mirror.ids)
    assert inc.shape == fresh.shape == (n, mirror.rows.shape[1])
    rows, ids = inc.read_rows(0, n)
    assert ids.tolist() == mirror.ids.tolist() and rows.tobytes() == mirror.rows.tobytes()
    for q in qs:
        for k in ks:
            assert _same(inc.query(q, k), fresh.query(q, k)), f"k={k}"
    if batch:
        s1, i1, c1 = inc.query_batch(qs, ks[0])
        s2, i2, c2 = fresh.query_batch(qs, ks[0])
        assert np.array_equal(c1, c2) and np.array_equal(i1, i2) and np.array_equal(s1.view(np.uint32), s2.view(np.uint32))


@pytest.mark.parametrize("devices,d,normalize", [([0], 96, False), ([0], 97, False), ([0, 0, 0], 64, False), ([0], 48, True)])
def test_thousand_random_mutations_equal_a_fresh_rebuild(devices, d, normalize, monkeypatch):
    import svs_b200
    monkeypatch.setenv("SVSB_ALLOW_DUP_DEVICES", "1")
    rng = np.random.default_rng(len(devices) * 1000 + d)
    n0 = 2500
    gen_rows = (lambda c: oracle.synth_matrix_normal(c, d, int(rng.integers(1 << 30)))) if not normalize else \
        (lambda c: rng.standard_normal((c, d)).astype(np.float32) * 3.0)
    store = lambda r: r if not normalize else None
    m0 = gen_rows(n0)
    ids0 = np.arange(1, n0 + 1, dtype=np.int64)
    qs = oracle.synth_queries(3, d, 5, dist="normal")
    inc, fresh = svs_b200.Engine(devices), svs_b200.Engine([0])
    try:
        inc.load(m0, ids0, normalize=normalize)
        if normalize:                                              # the mirror holds what the device stores
            m0 = inc.read_rows(0, n0)[0]
        mirror = Mirror(m0, ids0)
        steps = 1000 if len(devices) == 1 and d == 96 else 250
        snap, snap_state = None, None
        for step in range(steps):
            op = rng.choice(["add", "del", "both", "recycle"], p=[0.4, 0.3, 0.2, 0.1])
            dels, add_rows, add_ids = np.zeros(0, np.int64), None, np.zeros(0, np.int64)
            if op in ("del", "both") and len(mirror.ids) > 50:
                dels = rng.choice(mirror.ids, size=int(rng.integers(1, 4)), replace=False)
            if op == "recycle" and len(mirror.ids) > 50:
                dels = mirror.ids[-1:].copy()                      # delete the newest row: its id comes back
            mirror.delete(dels)
            if op in ("add", "both", "recycle"):
                c = int(rng.integers(1, 6))
                raw = gen_rows(c)
                add_ids = np.arange(mirror.next_id(), mirror.next_id() + c, dtype=np.int64)
                add_rows = raw
            inc.apply_mutations(dels, add_ids, add_rows)
            if add_rows is not None:
                stored = add_rows if not normalize else inc.read_rows(len(mirror.ids), len(add_ids))[0]
                mirror.append(stored, add_ids)
            if step % 10 == 0 or step == steps - 1:
                _compare(inc, fresh, mirror, qs, (7, 100) if step % 50 else (7, 100, 2048, 2300, 10**9), batch=step % 50 == 0)
            if step == 100:                                        # a snapshot pins the generation it was taken on
                snap, snap_state = inc.snapshot(), (mirror.rows.copy(), mirror.ids.copy())
        if snap is not None:
            fresh.load(*snap_state)
            for q in qs:
                assert _same(snap.query(q, 60), fresh.query(q, 60))
            snap.release()
        phys, live = inc.generation_rows()
        assert live == len(mirror.ids) and phys > live               # tombstones are really there
        if not normalize:
            dev, bad = inc.norm_stats()
            assert bad == 0
    finally:
        inc.close(); fresh.close()


def test_refusals_leave_the_resident_generation_untouched():
    import svs_b200
    from svs_b200 import _lib
    d = 32
    m = oracle.synth_matrix_normal(500, d, 3)
    ids = np.arange(10, 510, dtype=np.int64)
    q = oracle.synth_queries(1, d, 4, dist="normal")[0]
    with svs_b200.Engine([0]) as e:
        with pytest.raises(svs_b200.EngineError) as ex:
            e.apply_mutations([10], [], None)                      # nothing resident
        assert ex.value.code == _lib.SVSB_E_STATE
        e.load(m, ids)
        before = e.query(q, 20)
        row = oracle.synth_matrix_normal(2, d, 9)
        for dels, aids, rows in [([9999], [], None),               # not a row of the matrix
                                 ([10, 10], [], None),             # the same row twice
                                 ([], [600, 600], row),            # ids not ascending
                                 ([], [509, 700], row),            # 509 is a live id
                                 ([], [700, 701], oracle.synth_matrix_normal(2, d + 1, 9))]:   # wrong row length
            with pytest.raises(svs_b200.EngineError) as ex:
                e.apply_mutations(dels, aids, rows)
            assert ex.value.code == _lib.SVSB_E_STATE
            assert _same(e.query(q, 20), before) and e.generation_rows() == (500, 500)
        e.apply_mutations([509], [509, 700], row)                  # after deleting 509 its id may come back
        assert e.generation_rows() == (502, 501)
        got = e.retrieve(row[0], 1)
        assert got[0][1] == 509 and got[0][0] == pytest.approx(1.0, abs=1e-5)
        # every row deleted: an empty matrix, as np.dot sees it after the reference rebuilt (0, 0)
        e.apply_mutations(np.concatenate([np.arange(10, 509), [509, 700]]), [], None)
        assert e.shape[0] == 0
        with pytest.raises(ValueError):
            e.query(q, 5)


def test_appends_grow_the_buffer_geometrically_and_keep_old_generations_alive():
    import svs_b200
    d = 128
    m = oracle.synth_matrix_normal(3000, d, 11)
    ids = np.arange(1, 3001, dtype=np.int64)
    q = oracle.synth_queries(1, d, 12, dist="normal")[0]
    with svs_b200.Engine([0]) as e, svs_b200.Engine([0]) as fresh:
        e.load(m, ids)
        snaps = [e.snapshot()]
        want0 = e.query(q, 10)
        rows_all, ids_all = m, ids
        for step in range(12):                                     # 12 x 700 rows: several re-allocations
            add = oracle.synth_matrix_normal(700, d, 100 + step)
            aid = np.arange(ids_all[-1] + 1, ids_all[-1] + 701, dtype=np.int64)
            e.apply_mutations([], aid, add)
            rows_all, ids_all = np.concatenate([rows_all, add]), np.concatenate([ids_all, aid])
            if step % 4 == 0:
                snaps.append(e.snapshot())
        fresh.load(rows_all, ids_all)
        assert _same(e.query(q, 300), fresh.query(q, 300))
        s, i, c = e.query_batch(np.stack([q] * 9), 50)             # no tombstones: the batched path may use its tensor-core pass
        assert _same((s[0], i[0]), fresh.query(q, 50))
        assert _same(snaps[0].query(q, 10), want0)                 # the first generation still answers from ITS rows
        for sn in snaps:
            sn.release()
