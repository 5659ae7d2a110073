"""BASELINE.json's full taxi size (1.1 B rows on one GPU) through size-independent properties — the oracle cannot run
this size in seconds, so parity here is: conservation of rows, marginals against independent torch reductions of the
same columns, agreement between queries that share keys, idempotence, and split invariance (whole table ==
device-side reduce of two halves).  HDK_B200_FULLSIZE_ROWS overrides the row count."""
import ctypes as C
import os

import numpy as np
import pytest

from hdk_b200 import abi
from tests import util

pytestmark = pytest.mark.gpu
ROWS = int(os.environ.get("HDK_B200_FULLSIZE_ROWS", 1_100_000_000))


@pytest.fixture(scope="module")
def taxi(oracle_mod):
    import torch
    import benchdata
    from hdk_b200.executor import Executor
    from hdk_b200.storage import ArrowStorage
    free, _ = torch.cuda.mem_get_info()
    rows = ROWS if free > ROWS * 34 else max(int(free // 48), 1_000_000)   # 30 B/row of columns + head room
    st = ArrowStorage()
    benchdata.make_taxi(st, torch.device("cuda", 0), rows)
    return st, Executor(st), rows


def _run(ex, st, text, fragments=None):
    import torch
    from hdk_b200 import sql
    pq = ex.plan(sql.parse(text.split(" ORDER BY")[0], st.tables))
    prep = ex.prepare(pq, fragments=fragments)
    ex.launch(pq, prep)
    torch.cuda.synchronize()
    assert int(prep["err"].item()) == 0
    return pq, prep


def _rows(oracle_mod, pq, buf, nk):
    return util.sort_rows(util.result_columns(oracle_mod, pq, buf), nk)


def test_full_size_taxi_properties(oracle_mod, taxi):
    import torch
    import benchdata
    st, ex, n = taxi
    frags = st.get_table("trips").fragments
    col = lambda name: [f.device_chunks[name] for f in frags]   # noqa: E731
    # Q1: counts per cab_type == bincount of the column; Σ == N
    pq1, p1 = _run(ex, st, benchdata.TAXI_QUERIES["q1"])
    r1 = _rows(oracle_mod, pq1, p1["out"].cpu().numpy(), 1)
    ones = sum(int(c.view(torch.int32).sum().item()) for c in col("cab_type"))
    assert dict((k, c) for k, c in r1) == {0: n - ones, 1: ones}
    # Q2: AVG(total_amount) per passenger_count vs torch (independent fp64 reduction order), 1e-9 relative
    pq2, p2 = _run(ex, st, benchdata.TAXI_QUERIES["q2"])
    r2 = _rows(oracle_mod, pq2, p2["out"].cpu().numpy(), 1)
    sums, cnts = torch.zeros(10, dtype=torch.float64, device="cuda"), torch.zeros(10, dtype=torch.int64, device="cuda")
    for pc, ta in zip(col("passenger_count"), col("total_amount")):
        k = pc.view(torch.int16).to(torch.int64)
        sums.index_add_(0, k, ta.view(torch.float64))
        cnts += torch.bincount(k, minlength=10)
    assert int(cnts.sum().item()) == n
    exp2 = [(int(i), float(sums[i] / cnts[i])) for i in range(10) if int(cnts[i]) > 0]
    util.assert_rows_equal(r2, exp2, rel=1e-9)
    # Q3: Σ counts == N and the passenger_count marginal == bincount
    pq3, p3 = _run(ex, st, benchdata.TAXI_QUERIES["q3"])
    r3 = _rows(oracle_mod, pq3, p3["out"].cpu().numpy(), 2)
    assert sum(r[2] for r in r3) == n
    marg = {}
    for pc, y, c in r3:
        marg[pc] = marg.get(pc, 0) + c
    assert marg == {int(i): int(cnts[i]) for i in range(10) if int(cnts[i]) > 0}
    assert {r[1] for r in r3} <= set(range(2009, 2017))
    # Q4: Σ counts == N and its (passenger_count, year) marginal == Q3
    pq4, p4 = _run(ex, st, benchdata.TAXI_QUERIES["q4"])
    r4 = _rows(oracle_mod, pq4, p4["out"].cpu().numpy(), 3)
    assert sum(r[3] for r in r4) == n
    m4 = {}
    for pc, y, d, c in r4:
        m4[(pc, y)] = m4.get((pc, y), 0) + c
    assert m4 == {(pc, y): c for pc, y, c in r3}
    assert all(0 <= r[2] <= 200 for r in r4)
    # idempotence: integer results are bit-identical run to run
    for q, (pq, p) in {"q1": (pq1, p1), "q3": (pq3, p3), "q4": (pq4, p4)}.items():
        first = p["out"].clone()
        ex.launch(pq, p)
        torch.cuda.synchronize()
        assert torch.equal(first, p["out"]), q
    # split invariance: whole table == hdk_b200_reduce(first half, second half)
    L = ex.lib
    half = len(frags) // 2
    if half:
        for q in ("q3", "q4"):
            pq, whole = {"q3": (pq3, p3), "q4": (pq4, p4)}[q]
            bufs = []
            for part in (frags[:half], frags[half:]):
                _, pp = _run(ex, st, benchdata.TAXI_QUERIES[q], fragments=part)
                bufs.append(pp["out"].clone())
            err = torch.zeros(1, dtype=torch.int32, device="cuda")
            assert L.hdk_b200_reduce(C.byref(pq.plan), C.byref(pq.qmd), bufs[0].data_ptr(), bufs[1].data_ptr(), pq.qmd.entry_count,
                                     err.data_ptr(), None) == 0
            torch.cuda.synchronize()
            assert int(err.item()) == 0 and torch.equal(bufs[0], whole["out"]), q
