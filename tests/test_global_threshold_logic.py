"""The LOGIC of the global-threshold batch protocol, exercised on adversarial data with the NumPy
restatement of the engine's kernels (tests/test_sharded_gloo.py::NumpyGlobalBatchBackend) and no process group: whatever the
data -- every top-k row in one shard, rows stored in score order, heavy duplicates, a sample that misses the top entirely, a
record capacity far too small -- a query must come out either EXACTLY as the reference's superheavy() ranks it or as -1 ("redo
with the exact path"); never a wrong list."""
import numpy as np
import pytest
import torch

from _util import oracle
from test_sharded_gloo import NumpyGlobalBatchBackend

from svs_b200.sharded import partition


def _run(m, ids, qs, world, k, sample_rank, cap):
    n = len(m)
    backs = []
    for r in range(world):
        b = NumpyGlobalBatchBackend()
        row0, cnt = partition(n, world, r)
        b.set_shard(row0)
        b.load_rows(m[row0:row0 + cnt], ids[row0:row0 + cnt])
        backs.append(b)
    probes = [b.batch_global_probe(k) for b in backs]
    if not all(p[0] for p in probes):
        return None
    norm = max(p[3] for p in probes)
    dq = torch.from_numpy(qs.copy())
    tops = [b.new_tops(len(qs)) for b in backs]
    for r, b in enumerate(backs):
        b.batch_sample_tops(dq, k, norm, tops[r])
    tops_all = torch.stack(tops, dim=0)
    recs = [b.new_records(len(qs), cap) for b in backs]
    for r, b in enumerate(backs):
        b.batch_global_records(dq, k, tops_all, world, sample_rank, cap, recs[r])
    o_s, o_i, o_c = backs[0].new_outputs(len(qs), k)
    backs[0].enqueue_merge_verified(torch.stack(recs, dim=0), world, len(qs), cap, k, min(k, n), o_s, o_i, o_c)
    return o_s.numpy(), o_i.numpy(), o_c.numpy()


def _check(m, ids, qs, world, k, sample_rank, cap, want_some_answers=True):
    out = _run(m, ids, qs, world, k, sample_rank, cap)
    assert out is not None
    s, i, c = out
    answered = 0
    for j, q in enumerate(qs):
        if c[j] < 0:
            continue
        answered += 1
        want = oracle.superheavy(m, ids, q, k)
        got = list(zip(s[j, :c[j]].tolist(), i[j, :c[j]].tolist()))
        oracle.compare_retrieval(got, want, oracle.scores_of(m, q), ids)
    if want_some_answers:
        assert answered > 0
    return answered


CASES = ["random", "all_top_in_one_shard", "rows_sorted_by_score", "duplicates", "clustered"]


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("world", [2, 5])
def test_a_query_is_answered_exactly_or_refused(case, world):
    rng = np.random.default_rng(abs(hash((case, world))) % (2 ** 32))
    n, d, k, b = 3000, 24, 20, 12
    m = oracle.synth_matrix_normal(n, d, 7)
    qs = oracle.synth_queries(b, d, 8, "normal")
    if case == "all_top_in_one_shard":          # the k best rows of every query live in the LAST shard
        base = qs.mean(axis=0); base /= np.linalg.norm(base)
        near = base[None, :] + 0.05 * rng.standard_normal((2 * k, d)).astype(np.float32)
        m[-2 * k:] = near / np.linalg.norm(near, axis=1)[:, None]
        qs = (base[None, :] + 0.02 * rng.standard_normal((b, d))).astype(np.float32)
        qs /= np.linalg.norm(qs, axis=1)[:, None]
    elif case == "rows_sorted_by_score":        # storage order correlated with the first query: strided samples are NOT random
        order = np.argsort(-(m @ qs[0]))
        m = np.ascontiguousarray(m[order])
    elif case == "duplicates":                  # 40 % of the rows are copies of 30 originals: massive exact ties
        src = rng.integers(0, 30, size=int(0.4 * n))
        m[rng.choice(n, size=len(src), replace=False)] = m[src]
    elif case == "clustered":                   # scores packed in a band narrower than 2 eps
        centre = oracle.synth_queries(1, d, 9, "normal")[0]
        m = (centre[None, :] + 2e-4 * rng.standard_normal((n, d))).astype(np.float32)
        m /= np.linalg.norm(m, axis=1)[:, None]
    m = np.ascontiguousarray(m, dtype=np.float32)
    ids = np.cumsum(rng.integers(1, 4, size=n)).astype(np.int64)
    f = NumpyGlobalBatchBackend.SAMPLE_STRIDE
    lam = k / f
    rank = int(np.ceil(lam + 6.0 * np.sqrt(lam) + 4.0))
    share = k / world
    cap = min(k, int(np.ceil(share + 6.0 * np.sqrt(share) + 4.0)))
    # the planned parameters: exact or refused; random data must be mostly answered
    answered = _check(m, ids, qs, world, k, min(rank, 32), cap, want_some_answers=(case == "random"))
    if case == "random":
        assert answered >= b - 1
    # hostile parameters: an order statistic far too high (rank 1), a record capacity far too small (1, 2)
    _check(m, ids, qs, world, k, 1, cap, want_some_answers=False)
    _check(m, ids, qs, world, k, min(rank, 32), 1, want_some_answers=False)
    _check(m, ids, qs, world, k, 32, 2, want_some_answers=False)


def test_k_larger_than_a_shard_and_than_the_matrix():
    n, d = 1600, 16
    m = oracle.synth_matrix_normal(n, d, 17)
    ids = np.arange(5, n + 5, dtype=np.int64)
    qs = oracle.synth_queries(4, d, 18, "normal")
    for k in (600, 1600):                       # a 3-rank shard has 534 rows
        _check(m, ids, qs, 3, k, 32, k, want_some_answers=False)
