"""GPU: several devices in ONE process (svsb_create with n_dev > 1; multi.cu) -- what `install(svs, devices=[...])` puts
behind KB.retrieve.  On a one-GPU box the devices are virtual shards of device 0 (SVSB_ALLOW_DUP_DEVICES=1); with more
GPUs visible the same tests use them.  Everything must equal the single-device engine bit for bit: the fused path
(selection kernels push their records into device 0's gather window, a merge kernel waits on the flags), k > 2048
(peer copies + rank merge), batches, submit / wait with several queries in flight, caller threads."""
import os
import threading

import numpy as np
import pytest

from _util import oracle

pytestmark = pytest.mark.gpu


def _devices(count):
    try:
        import torch
        have = torch.cuda.device_count()
    except Exception:
        have = 1
    return [i % max(have, 1) for i in range(count)]


def _same(a, b):
    return len(a[0]) == len(b[0]) and np.array_equal(a[0].view(np.uint32), b[0].view(np.uint32)) and np.array_equal(a[1], b[1])


@pytest.fixture(scope="module")
def engines():
    import svs_b200
    os.environ["SVSB_ALLOW_DUP_DEVICES"] = "1"
    one = svs_b200.Engine([0])
    multi = {n: svs_b200.Engine(_devices(n)) for n in (2, 3, 8)}
    yield one, multi
    one.close()
    for e in multi.values():
        e.close()


def test_fused_path_large_k_and_batches_equal_the_single_device_engine(engines):
    one, multi = engines
    n, d = 50_003, 192
    m = oracle.synth_matrix_uniform(n, d, 21)
    m[7] = m[n - 3]                                                # an exact tie across the first and the last shard
    ids = np.cumsum(np.random.default_rng(1).integers(1, 3, size=n)).astype(np.int64)
    qs = oracle.synth_queries(6, d, 22)
    qs[2] = m[7]
    one.load(m, ids)
    oracle.compare_retrieval(one.retrieve(qs[0], 100), oracle.superheavy(m, ids, qs[0], 100), oracle.scores_of(m, qs[0]), ids)
    for nd, e in multi.items():
        e.load(m, ids)
        back, bid = e.read_rows(0, n)
        assert back.tobytes() == m.tobytes() and (bid == ids).all()
        for q in qs:
            for k in (1, 10, 100, 2048, 2049, 5000, n, 10**9):     # the reference answers any n (util.py:198-199)
                assert _same(e.query(q, k), one.query(q, k)), (nd, k)
        assert [i for _, i in e.retrieve(qs[2], 2)] == [int(ids[7]), int(ids[n - 3])]
        for k in (5, 100, 3000):
            s1, i1, c1 = e.query_batch(qs, k)
            s2, i2, c2 = one.query_batch(qs, k)
            assert np.array_equal(c1, c2) and np.array_equal(i1, i2) and np.array_equal(s1.view(np.uint32), s2.view(np.uint32)), (nd, k)
        big = oracle.synth_queries(300, d, 23)                     # large batch: every shard takes its tensor-core path
        s1, i1, c1 = e.query_batch(big, 20)
        s2, i2, c2 = one.query_batch(big, 20)
        assert np.array_equal(c1, c2) and np.array_equal(i1, i2) and np.array_equal(s1.view(np.uint32), s2.view(np.uint32)), nd
        assert e.retrieve(qs[0], 0) == [] and e.query_batch(qs, 0)[2].tolist() == [0] * len(qs)
        with pytest.raises(ValueError):
            e.query(np.zeros(d + 1, np.float32), 3)


def test_fewer_rows_than_devices_and_reloads(engines):
    one, multi = engines
    d = 16
    m = oracle.synth_matrix_normal(5, d, 2)
    ids = np.array([3, 4, 9, 10, 11], dtype=np.int64)
    q = oracle.synth_queries(1, d, 3, dist="normal")[0]
    one.load(m, ids)
    for nd, e in multi.items():
        for rows in (5, 2, 1):
            one.load(m[:rows], ids[:rows]); e.load(m[:rows], ids[:rows])
            for k in (1, 3, 50, 5000):
                assert _same(e.query(q, k), one.query(q, k)), (nd, rows, k)
            s1, i1, c1 = e.query_batch(np.stack([q, q, q]), 4)
            assert c1.tolist() == [min(4, rows)] * 3
        e.load(np.zeros((0, 0), np.float32), np.zeros(0, np.int64))
        assert e.shape == (0, 0)
        with pytest.raises(ValueError):
            e.query(q, 1)


def test_submit_wait_keeps_several_queries_in_flight(engines):
    one, multi = engines
    n, d, k = 30_000, 256, 50
    m = oracle.synth_matrix_uniform(n, d, 31)
    ids = np.arange(1, n + 1, dtype=np.int64)
    qs = oracle.synth_queries(40, d, 32)
    one.load(m, ids)
    want = [one.query(q, k) for q in qs]
    for e in [one] + list(multi.values()):
        e.load(m, ids)
        pend = []
        got = []
        for q in qs:                                               # at most 3 pending: submit, then drain the oldest
            pend.append(e.submit(q, k))
            if len(pend) == 3:
                got.append(pend.pop(0).result())
        got += [p.result() for p in pend]
        assert all(_same(a, b) for a, b in zip(got, want))
        assert len(e.submit(qs[0], 0).result()[0]) == 0            # k <= 0: nothing to compute, still a handle
        with pytest.raises(ValueError):
            e.submit(np.zeros(d + 3, np.float32), k)
        # a caller that holds every slot as a pending handle gets an error, not a wait for itself
        import svs_b200
        os.environ["SVSB_SUBMIT_TIMEOUT_MS"] = "200"
        held = []
        with pytest.raises(svs_b200.EngineError, match="pending already"):
            for _ in range(8):
                held.append(e.submit(qs[1], k))
        assert len(held) in (3, 4) and all(_same(p.result(), want[1]) for p in held)
        del os.environ["SVSB_SUBMIT_TIMEOUT_MS"]


def test_a_pending_query_does_not_block_loads_or_other_query_forms_on_its_thread(engines):
    """One thread: submit, then -- with the handle still pending -- a full ranking, a batch and a reload.  None of them may
    wait for the handle to be waited for (they would wait for ever)."""
    one, multi = engines
    n, d = 12_000, 64
    m = oracle.synth_matrix_normal(n, d, 61)
    ids = np.arange(1, n + 1, dtype=np.int64)
    qs = oracle.synth_queries(4, d, 62, dist="normal")
    one.load(m, ids)
    want = one.query(qs[0], 30)
    for e in multi.values():
        e.load(m, ids)
        pend = [e.submit(qs[0], 30), e.submit(qs[0], 30)]
        assert _same(e.query(qs[1], n), one.query(qs[1], n))              # k > 2048: the gather path
        s, i, c = e.query_batch(qs, 9)
        assert _same((s[2], i[2]), one.query(qs[2], 9))
        e.load(m[:6000], ids[:6000])                                       # a new generation of a different size
        assert all(_same(p.result(), want) for p in pend)                  # the pending queries answer from THEIR generation
        one.load(m[:6000], ids[:6000])
        assert _same(e.query(qs[0], 30), one.query(qs[0], 30))
        one.load(m, ids)


def test_caller_threads_share_a_multi_device_engine(engines):
    one, multi = engines
    n, d = 20_000, 128
    m = oracle.synth_matrix_normal(n, d, 41)
    ids = np.arange(100, 100 + n, dtype=np.int64)
    qs = oracle.synth_queries(16, d, 42, dist="normal")
    one.load(m, ids)
    want = {k: [one.query(q, k) for q in qs] for k in (10, 100)}
    e = multi[3]
    e.load(m, ids)
    errors = []

    def worker(t):
        try:
            for rep in range(25):
                j = (t * 7 + rep) % len(qs)
                k = 10 if (t + rep) % 2 else 100
                if rep % 5 == 4:
                    s, i, c = e.query_batch(qs[j:j + 2], k)
                    assert _same((s[0], i[0]), want[k][j])
                else:
                    assert _same(e.query(qs[j], k), want[k][j])
        except Exception as ex:                                    # noqa: BLE001
            errors.append(repr(ex))
    threads = [threading.Thread(target=worker, args=(t,)) for t in range(6)]
    for t in threads:
        t.start()
    for _ in range(3):
        e.load(m, ids)                                             # reloads race with the queries: same data, same answers
    for t in threads:
        t.join()
    assert errors == []


def test_bench_loop_on_a_multi_device_engine(engines):
    one, multi = engines
    n, d, k = 40_000, 512, 100
    m = oracle.synth_matrix_uniform(n, d, 51)
    ids = np.arange(1, n + 1, dtype=np.int64)
    qs = oracle.synth_queries(5, d, 52)
    one.load(m, ids)
    for nd, e in multi.items():
        e.load(m, ids)
        e.bench_set_queries(qs)
        r = e.bench_run(k, 13, with_gemv=True)
        assert r["total_ms"] > 0 and r["gemv_ms"] > 0 and r["launches"] >= 13 * (2 * nd + 1)
        got = e.bench_last_result(k)                               # the 13th query = qs[12 % 5]
        assert got == one.retrieve(qs[12 % 5], k)
