"""The N>1 path on CPU: world_size-2 (and 3) `gloo` runs of ShardedRetriever with a NumPy stand-in for the
shard compute (the CUDA backend is covered by the GPU suite).  Exercises the partitioning, the packed
record layout, the all-gather, micro-batching and the merge order -- the host logic of SURVEY 8e."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from _util import oracle

from svs_b200.sharded import MICRO_BATCH, ShardedRetriever, partition


def np_keys(scores: np.ndarray, rows: np.ndarray) -> np.ndarray:
    """The engine's 64-bit key (svs_b200/csrc/common.cuh make_key), restated in NumPy."""
    bits = scores.astype(np.float32).view(np.uint32)
    ordered = np.where(bits & np.uint32(0x80000000), ~bits, bits | np.uint32(0x80000000)).astype(np.uint64)
    return (ordered << np.uint64(32)) | (~rows.astype(np.uint32)).astype(np.uint64)


def key_score(keys: np.ndarray) -> np.ndarray:
    o = (keys >> np.uint64(32)).astype(np.uint32)
    bits = np.where(o & np.uint32(0x80000000), o & np.uint32(0x7FFFFFFF), ~o)
    return bits.astype(np.uint32).view(np.float32)


class NumpyShardBackend:
    """CPU stand-in with the interface of svs_b200.sharded.CudaShardBackend (test infrastructure)."""

    def __init__(self):
        self.row0 = 0

    def set_shard(self, row0):
        self.row0 = row0

    def load_rows(self, rows, ids):
        self.m, self.ids = rows, ids

    def device_queries(self, Q):
        return torch.from_numpy(np.ascontiguousarray(Q, dtype=np.float32).copy())

    def new_records(self, count, k):
        return torch.zeros((count, 2 * k + 1), dtype=torch.int64)

    def new_outputs(self, count, k):
        return (torch.zeros((count, k), dtype=torch.float32), torch.zeros((count, k), dtype=torch.int64),
                torch.zeros((count,), dtype=torch.int32))

    def join(self):
        pass

    def batch_local(self, queries, k, records):
        for j in range(queries.shape[0]):
            self.enqueue_local(queries[j], k, records[j])
        return 0

    def enqueue_local(self, q, k, record_row, time_kernel=False, seq=0):
        x = oracle.scores_of(self.m, q.numpy()) if len(self.m) else np.zeros(0, np.float32)
        keys = np_keys(x, np.arange(self.row0, self.row0 + len(x)))
        order = np.argsort(keys)[::-1][:k]
        rec = record_row.numpy()
        rec[:len(order)] = keys[order].view(np.int64)
        rec[k:k + len(order)] = self.ids[order]
        rec[2 * k] = len(order)

    def enqueue_merge(self, gathered, n_lists, batch, k, out_scores, out_ids, out_counts):
        g = gathered.numpy().reshape(n_lists, batch, 2 * k + 1)
        for b in range(batch):
            keys, ids = [], []
            for l in range(n_lists):
                c = int(g[l, b, 2 * k] & 0xFFFFFFFF)
                keys.append(g[l, b, :c].view(np.uint64)); ids.append(g[l, b, k:k + c])
            keys, ids = np.concatenate(keys), np.concatenate(ids)
            order = np.argsort(keys)[::-1][:k]
            out_scores.numpy()[b, :len(order)] = key_score(keys[order])
            out_ids.numpy()[b, :len(order)] = ids[order]
            out_counts.numpy()[b] = len(order)

    def collect_kernel_ms(self):
        return 0.0

    def close(self):
        pass


class NumpyPeerBackend(NumpyShardBackend):
    """CPU restatement of the fused peer exchange (include/svsb200.h "peer exchange"; csrc/select.cu peer_publish /
    merge_window_kernel): every rank owns a gather window -- here a POSIX shared-memory block instead of HBM opened
    over CUDA IPC -- of SLOTS x world records [keys(cap) | ids(cap) | count | pad] plus SLOTS x world flag words.
    Query number seq uses slot seq % SLOTS; a rank stores its record into EVERY rank's window, then the flag = seq;
    the merge spins on its own window's flags.  Test infrastructure for ShardedRetriever's exchange="peer" host
    logic (handle all-gather, connect, SPMD sequence numbers, slot reuse)."""
    SLOTS = 4

    def exchange_handle(self, world, rank, k_max=64):
        from multiprocessing import shared_memory
        self.world, self.rank, self.cap, self.seq = world, rank, k_max, 0
        self.rec_words = 2 * k_max + 2
        words = self.SLOTS * world + self.SLOTS * world * self.rec_words
        self._own = shared_memory.SharedMemory(create=True, size=words * 8)
        np.ndarray((words,), dtype=np.uint64, buffer=self._own.buf)[:] = 0
        return self._own.name.encode().ljust(64, b"\0")            # 64 bytes, like a cudaIpcMemHandle_t

    def exchange_connect(self, handles):
        from multiprocessing import shared_memory
        assert len(handles) == self.world and all(len(h) == 64 for h in handles)
        self._shm, self._win = [], []
        for r, h in enumerate(handles):
            shm = self._own if r == self.rank else shared_memory.SharedMemory(name=h.rstrip(b"\0").decode())
            self._shm.append(shm)
            self._win.append(np.ndarray((shm.size // 8,), dtype=np.uint64, buffer=shm.buf))

    def _rec(self, win, slot, src):
        base = self.SLOTS * self.world + (slot * self.world + src) * self.rec_words
        return win[base:base + self.rec_words]

    def _query(self, q, k):
        import time
        self.seq += 1
        slot = self.seq % self.SLOTS
        tmp = torch.zeros(2 * k + 1, dtype=torch.int64)
        self.enqueue_local(q, k, tmp)                              # local record, the collective path's layout
        t = tmp.numpy().view(np.uint64)
        cnt = int(t[2 * k])
        for p in range(self.world):                                # the "epilogue": push into every window, then publish
            rec = self._rec(self._win[p], slot, self.rank)
            rec[:cnt] = t[:cnt]
            rec[self.cap:self.cap + cnt] = t[k:k + cnt]
            rec[2 * self.cap] = cnt
            self._win[p][slot * self.world + self.rank] = self.seq
        own = self._win[self.rank]
        deadline = time.time() + 30
        while not all(int(own[slot * self.world + r]) >= self.seq for r in range(self.world)):   # the waiting merge
            assert time.time() < deadline, "a peer's record did not arrive"
            time.sleep(0.0002)
        keys, ids = [], []
        for r in range(self.world):
            rec = self._rec(own, slot, r)
            c = int(rec[2 * self.cap])
            keys.append(rec[:c].copy()); ids.append(rec[self.cap:self.cap + c].copy().view(np.int64))
        keys, ids = np.concatenate(keys), np.concatenate(ids)
        order = np.argsort(keys)[::-1][:k]
        return key_score(keys[order]), ids[order]

    def query_peer(self, q, k):
        return self._query(torch.from_numpy(np.ascontiguousarray(q, dtype=np.float32)), k)

    def enqueue_query_peer(self, q, k, out_scores, out_ids, out_count, time_kernel=False, pipelined=True):
        s, i = self._query(q, k)
        out_scores.numpy()[:len(s)] = s
        out_ids.numpy()[:len(i)] = i
        out_count.numpy()[...] = len(s)

    def close(self):
        for r, shm in enumerate(getattr(self, "_shm", [])):
            self._win[r] = None
        self._win = []
        for r, shm in enumerate(getattr(self, "_shm", [])):
            shm.close()
        if getattr(self, "_own", None) is not None:
            self._own.unlink()
            self._own = None


def _peer_worker(rank, world, port, n, d, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        m = oracle.synth_matrix_normal(n, d, 13)
        if n > 10:
            m[5] = m[n - 2]
        ids = np.cumsum(np.random.default_rng(1).integers(1, 4, size=n)).astype(np.int64)
        sr = ShardedRetriever(rank, world, backend=NumpyPeerBackend())
        assert sr.exchange == "peer"
        sr.load_global(m, ids)
        qs = oracle.synth_queries(MICRO_BATCH + 5, d, 14, "normal")
        if n > 10:
            qs[1] = m[5]
        out = []
        for k in (1, 10, 64):
            for q in qs[:6]:                                       # 18 queries: every window slot is reused 4 times
                got = sr.retrieve(q, k)
                oracle.compare_retrieval(got, oracle.superheavy(m, ids, q, k), oracle.scores_of(m, q), ids)
                out.append(got)
        if n > 10:
            assert [i for _, i in sr.retrieve(qs[1], 2)] == [int(ids[5]), int(ids[n - 2])]
        sr.set_queries(qs)
        assert sr.run_queries(10, len(qs)) == 0.0                 # device-resident loop: no collective inside
        s_, i_, c_ = sr._micro_batch([sr._queries[j] for j in range(3)], 10, False)
        for j in range(3):
            got = [(float(a), int(b)) for a, b in zip(s_[j].numpy(), i_[j].numpy())][:int(c_[j])]
            oracle.compare_retrieval(got, oracle.superheavy(m, ids, qs[j], 10), oracle.scores_of(m, qs[j]), ids)
            out.append(got)
        ret[rank] = out
        dist.barrier()
        sr.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 1001), (3, 2)])
def test_peer_exchange_protocol_on_cpu_matches_the_oracle_on_every_rank(world, n):
    """exchange="peer" end to end on CPU: handles all-gathered over gloo, windows connected, records pushed and
    flags awaited across PROCESSES (shared memory standing in for NVLink peer memory)."""
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_peer_worker, args=(world, _free_port(), n, 24, ret), nprocs=world, join=True)
    assert len(ret) == world
    for r in range(1, world):
        assert ret[r] == ret[0]


class NumpyGlobalBatchBackend(NumpyShardBackend):
    """CPU restatement of the global-threshold batch protocol (include/svsb200.h svsb_batch_global_*; csrc/batch.cu
    sample_order_kernel / union_threshold_kernel / refine_* with REFINE_PARTIAL | REFINE_DEFER; csrc/select.cu
    merge_records_verified): coarse scores = fp16 operands (scaled by 2^12) with an fp32 sum, eps as the engine bounds it,
    a strided sample, the 32 largest sample values per query exchanged, T = order statistic of the union - 2 eps,
    candidates = coarse >= T, all re-scored exactly, records [keys(cap) | ids(cap) | count (bit 30: truncated), ver],
    merge: any count < 0 or sum(ver) < verify_k or a truncated list whose last entry makes the top k -> count -1.
    Test infrastructure for ShardedRetriever._batch / _global_plan / the redo of unanswered queries."""
    TOPX = 32
    SAMPLE_STRIDE = 16
    spoil_odd_queries = False                                   # push T above everything for odd queries -> count -1

    def batch_global_probe(self, k):
        n = len(self.m)
        sample = len(range(0, n, self.SAMPLE_STRIDE))
        norm = float(np.sqrt((self.m.astype(np.float64) ** 2).sum(axis=1)).max()) * 1.000001 if n else 1.0
        return (n >= 64 and sample >= self.TOPX, sample, n, norm)

    def new_tops(self, count):
        return torch.zeros((count, self.TOPX), dtype=torch.float32)

    def _coarse(self, Q):
        m16 = (self.m * np.float32(4096.0)).astype(np.float16).astype(np.float32)
        q16 = (Q * np.float32(4096.0)).astype(np.float16).astype(np.float32)
        return (m16 @ q16.T).T * np.float32(2.0 ** -24)            # (b, n_local)

    def batch_sample_tops(self, queries, k, max_row_norm, tops):
        Q = queries.numpy()
        d = Q.shape[1]
        coef = 2.0 ** -10 + 2.0 ** -22 + d * (2.0 ** -22 + 2.0 ** -23)
        self._eps = (coef * np.sqrt((Q.astype(np.float64) ** 2).sum(axis=1)) * 1.000001 * max_row_norm + 1e-8).astype(np.float32)
        self._co = self._coarse(Q)
        samp = self._co[:, ::self.SAMPLE_STRIDE].astype(np.float16).astype(np.float32)
        tops.numpy()[:] = -np.sort(-samp, axis=1)[:, :self.TOPX]

    def batch_global_records(self, queries, k, tops_all, world, sample_rank, rec_cap, records):
        Q = queries.numpy()
        b = Q.shape[0]
        union = tops_all.numpy().reshape(world, -1, self.TOPX)[:, :b].transpose(1, 0, 2).reshape(b, -1)
        T = -np.sort(-union, axis=1)[:, sample_rank - 1] - 2 * self._eps
        rec = records.numpy()
        rec[:] = 0
        for j in range(b):
            t = T[j] + (1e3 if (self.spoil_odd_queries and j % 2) else 0.0)
            cand = np.nonzero(self._co[j] >= t)[0]
            ver = int((self._co[j][cand] >= np.float32(t) + 2 * self._eps[j]).sum())
            exact = oracle.scores_of(self.m, Q[j])[cand]         # the per-query stand-in's bits (same BLAS call)
            keys = np_keys(exact, cand + self.row0)
            order = np.argsort(keys)[::-1][:k]
            ship = order[:rec_cap]
            rec[j, :len(ship)] = keys[ship].view(np.int64)
            rec[j, rec_cap:rec_cap + len(ship)] = self.ids[cand[ship]]
            cnt = len(ship) | ((1 << 30) if len(order) > len(ship) else 0)
            rec[j, 2 * rec_cap] = np.int64(cnt) | (np.int64(ver) << 32)

    def enqueue_merge_verified(self, gathered, n_lists, batch, rec_cap, k, verify_k, out_scores, out_ids, out_counts):
        g = gathered.numpy().reshape(n_lists, batch, 2 * rec_cap + 1)
        for b in range(batch):
            keys, ids, last_of_truncated, ver, bad = [], [], [], 0, False
            for l in range(n_lists):
                word = int(g[l, b, 2 * rec_cap])
                c32 = np.int32(np.uint32(word & 0xFFFFFFFF))
                if c32 < 0:
                    bad = True
                    continue
                c = int(c32) & ~(1 << 30)
                ver += word >> 32
                kl = g[l, b, :c].view(np.uint64)
                keys.append(kl); ids.append(g[l, b, rec_cap:rec_cap + c])
                if int(c32) & (1 << 30):
                    last_of_truncated.append(kl[c - 1])
            keys = np.concatenate(keys) if keys else np.zeros(0, np.uint64)
            ids = np.concatenate(ids) if ids else np.zeros(0, np.int64)
            order = np.argsort(keys)[::-1]
            for last in last_of_truncated:                         # rank of a truncated list's last entry must be >= k
                bad = bad or int((keys > last).sum()) < k
            if bad or ver < verify_k:
                out_counts.numpy()[b] = -1
                continue
            order = order[:k]
            out_scores.numpy()[b, :len(order)] = key_score(keys[order])
            out_ids.numpy()[b, :len(order)] = ids[order]
            out_counts.numpy()[b] = len(order)


def _global_batch_worker(rank, world, port, n, d, k, spoil, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        m = oracle.synth_matrix_normal(n, d, 31)
        m[9] = m[n - 4]                                           # an exact tie across shards
        ids = np.cumsum(np.random.default_rng(5).integers(1, 4, size=n)).astype(np.int64)
        be = NumpyGlobalBatchBackend()
        be.spoil_odd_queries = spoil
        sr = ShardedRetriever(rank, world, backend=be)
        sr.load_global(m, ids)
        plan = sr._global_plan(k)
        assert plan is not None and 1 <= plan[0] <= 32 and 1 <= plan[2] <= k, plan
        assert sr._global_plan(400) is None      # the agreed order statistic would exceed the 32 exchanged maxima: per-rank thresholds
        qs = oracle.synth_queries(21, d, 32, "normal")
        qs[3] = m[9]
        many = sr.retrieve_many(qs, k)
        for j in range(len(qs)):
            oracle.compare_retrieval(many[j], oracle.superheavy(m, ids, qs[j], k), oracle.scores_of(m, qs[j]), ids)
            assert many[j] == sr.retrieve(qs[j], k)               # the per-query path: same bits
        assert [i for _, i in many[3][:2]] == [int(ids[9]), int(ids[n - 4])]
        # odd queries were refused by the merge (count -1 on every rank) and redone by the exact path
        assert sr.last_fallbacks == (len(qs) // 2 if spoil else 0), sr.last_fallbacks
        ret[rank] = (plan, many)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def _ineligible_worker(rank, world, port, n, d, k, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        m = oracle.synth_matrix_normal(n, d, 41)
        ids = np.arange(1, n + 1, dtype=np.int64)
        sr = ShardedRetriever(rank, world, backend=NumpyGlobalBatchBackend())
        sr.load_global(m, ids)
        # the last rank's shard is too small to take part (its probe says so): EVERY rank must fall back to per-rank thresholds
        assert sr._global_plan(k) is None
        qs = oracle.synth_queries(5, d, 42, "normal")
        many = sr.retrieve_many(qs, k)
        for j in range(len(qs)):
            oracle.compare_retrieval(many[j], oracle.superheavy(m, ids, qs[j], k), oracle.scores_of(m, qs[j]), ids)
        ret[rank] = many
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_global_threshold_plan_falls_back_on_every_rank_when_one_rank_cannot_take_part():
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_ineligible_worker, args=(3, _free_port(), 1489, 24, 10, ret), nprocs=3, join=True)   # shards of 497, 497, 495 rows: the last samples 31 < 32
    assert len(ret) == 3 and ret[1] == ret[0] and ret[2] == ret[0]


@pytest.mark.parametrize("world,n,k,spoil", [(2, 3001, 10, False), (3, 4000, 25, False), (2, 2500, 10, True)])
def test_global_threshold_batches_on_cpu_match_the_oracle_on_every_rank(world, n, k, spoil):
    """retrieve_many through the global-threshold protocol across PROCESSES: the probes' object all-gather picks one
    order statistic and one record capacity for all ranks, two tensor all-gathers per batch (sample maxima, records),
    the verifying merge, and the redo of refused queries through the per-query path."""
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_global_batch_worker, args=(world, _free_port(), n, 24, k, spoil, ret), nprocs=world, join=True)
    assert len(ret) == world
    for r in range(1, world):
        assert ret[r] == ret[0]


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n, d, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(0)
        m = oracle.synth_matrix_normal(n, d, 3)
        tie = n > 10
        if tie:
            m[5] = m[n - 2]                              # an exact tie across the two ends (different shards)
        ids = np.cumsum(rng.integers(1, 4, size=n)).astype(np.int64)
        sr = ShardedRetriever(rank, world, backend=NumpyShardBackend())
        sr.load_global(m, ids)
        assert (sr.row0, sr.local_rows) == partition(n, world, rank)
        qs = oracle.synth_queries(2 * MICRO_BATCH + 3, d, 4, "normal")
        if tie:
            qs[1] = m[5]                                 # makes rows 5 and n-2 tie at the top
        out = []
        for k in (1, 10, 64):
            for q in qs[:4]:
                got = sr.retrieve(q, k)
                want = oracle.superheavy(m, ids, q, k)
                oracle.compare_retrieval(got, want, oracle.scores_of(m, q), ids)
                out.append(got)
        if tie:
            top = sr.retrieve(qs[1], 2)
            assert [i for _, i in top] == [int(ids[5]), int(ids[n - 2])]  # tie -> ascending id, across ranks
        # bench path: micro-batched, more queries than one batch, uneven tail
        sr.set_queries(qs)
        assert sr.run_queries(10, len(qs)) == 0.0
        s_, i_, c_ = sr._micro_batch([sr._queries[j] for j in range(3)], 10, False)
        for j in range(3):
            got = [(float(a), int(b)) for a, b in zip(s_[j].numpy(), i_[j].numpy())][:int(c_[j])]
            oracle.compare_retrieval(got, oracle.superheavy(m, ids, qs[j], 10), oracle.scores_of(m, qs[j]), ids)
        assert sr.retrieve(qs[0], 0) == []
        with pytest.raises(ValueError):
            sr.retrieve(np.zeros(d + 1, np.float32), 3)
        # batched form: one all-gather and one merge for the whole batch; result j == retrieve(qs[j])
        many = sr.retrieve_many(qs[:7], 10)
        for j in range(7):
            oracle.compare_retrieval(many[j], oracle.superheavy(m, ids, qs[j], 10), oracle.scores_of(m, qs[j]), ids)
            assert many[j] == sr.retrieve(qs[j], 10)
        assert sr.retrieve_many(qs[:2], 0) == [[], []]
        out.append(many)
        ret[rank] = out
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 1001), (3, 700), (2, 3)])
def test_sharded_retrieve_matches_the_oracle_on_every_rank(world, n):
    mgr = mp.Manager()
    ret = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n, 24, ret), nprocs=world, join=True)
    assert len(ret) == world
    for r in range(1, world):
        assert ret[r] == ret[0]                          # every rank holds the identical answer


def test_partition_covers_all_rows_contiguously():
    for n in (0, 1, 7, 1000, 1_000_000, 10_000_000):
        for world in (1, 2, 3, 4, 8):
            spans = [partition(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == n
            for (a, ca), (b, _cb) in zip(spans, spans[1:]):
                assert b == a + ca or (ca == 0 and b == a)
