"""CPU: (1) n is clipped to the row count BEFORE any buffer is sized by it or it crosses a 32-bit C argument
(the reference's get_top_k clips, src/svs/util.py:198-199, so retrieve(q, n=10**12) returns all N documents);
(2) bench.py's parity judge really rejects wrong answers; (3) both bench arms print the same `config`."""
import asyncio
import importlib.util
import os
import sys

import numpy as np
import pytest

from _util import ROOT, oracle

from svs_b200.engine import clamp_k
from svs_b200.dropin import _Coalescer


def test_clamp_k_follows_get_top_k():
    assert clamp_k(5, 10) == 5
    assert clamp_k(10**12, 10) == 10 and clamp_k(2**31, 7) == 7 and clamp_k(2**32, 7) == 7     # ctypes would wrap these
    assert clamp_k(0, 10) == 0 and clamp_k(-3, 10) == 0                                         # util.py:200-201
    assert clamp_k(5, 0) == 0
    assert clamp_k(2**40, 2**33) == 0x7fffffff


class _Matrix:
    """Stand-in for DeviceMatrix that records the k each engine call ran at."""

    def __init__(self, rows):
        self.shape = (rows, 4)
        self.ks = []

    def retrieve(self, q, n):
        self.ks.append(("one", n))
        return [(0.0, i) for i in range(min(max(n, 0), self.shape[0]))]

    def retrieve_many(self, qs, n):
        self.ks.append(("many", n))
        assert n <= self.shape[0], "the batch was sized by an unclipped n"
        return [[(0.0, i) for i in range(n)] for _ in qs]


def test_coalescer_clips_every_callers_n_to_the_row_count():
    m = _Matrix(rows=50)

    async def go():
        co = _Coalescer()
        v = np.zeros(4, np.float32)
        return await asyncio.gather(co.submit(m, v, 3), co.submit(m, v, 10**12), co.submit(m, v, 2**31), co.submit(m, v, 0))
    a, b, c, d = asyncio.run(go())
    assert len(a) == 3 and len(b) == 50 and len(c) == 50 and d == []
    assert ("many", 50) in m.ks                                    # one batch, at the clipped maximum


def _bench():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    argv = sys.argv
    sys.argv = ["bench.py"]
    try:
        spec.loader.exec_module(mod)
    finally:
        sys.argv = argv
    return mod


def test_bench_parity_judge_accepts_the_oracle_and_rejects_wrong_answers():
    bench = _bench()
    n, d, k = 4000, 32, 10
    m = oracle.synth_matrix_uniform(n, d, 3)
    ids = np.arange(1, n + 1, dtype=np.int64)
    qs = oracle.synth_queries(3, d, 4)

    class Eng:                                                     # read_rows as Engine has it
        def read_rows(self, a, cnt):
            return m[a:a + cnt], ids[a:a + cnt]
    good = [oracle.canonical_top_k(oracle.scores_of(m, q), ids, k) for q in qs]
    rep = bench.parity_single(Eng(), n, k, qs, good)
    assert rep["checked"] == 3 and rep["tolerance_ok"] == 3 and rep["max_rel_err"] <= 1e-5
    bad = [list(g) for g in good]
    bad[1][4] = (bad[1][4][0], int(ids[-1]))                       # a row that is not in the top-k
    with pytest.raises(AssertionError):
        bench.parity_single(Eng(), n, k, qs, bad)
    off = [list(g) for g in good]
    off[0][0] = (off[0][0][0] * 1.001, off[0][0][1])               # score off by 1e-3 relative
    with pytest.raises(AssertionError):
        bench.parity_single(Eng(), n, k, qs, off)


def test_both_bench_arms_print_the_same_config():
    bench = _bench()
    for name in bench.WORKLOADS:
        c = bench.config_of(name)
        assert set(c) == {"workload", "rows", "dims", "k", "l2"} and c == bench.config_of(name)
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert src.count('"config": config_of(') >= 5 and "NCCL_DEBUG" not in src
