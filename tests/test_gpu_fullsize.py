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


def test_reduce_and_launch_with_buffers_beyond_4_gib(oracle_mod):
    """ResultSetTest.cpp ReduceLargeBuffers.*Overflow32: group-by buffers larger than 2^32 bytes (150 M baseline entries
    x 32 B = 4.8 GB).  Launch, finalize, reduce and compaction must address them with 64-bit offsets: split invariance
    (whole == reduce of two halves) as decoded rows, entries must exist beyond the 4 GiB offset, and the decoded rows
    must equal the oracle's on a normally sized table."""
    import pyarrow as pa
    import torch
    from hdk_b200 import sql
    from hdk_b200.executor import Executor, ResultSet
    free, _ = torch.cuda.mem_get_info()
    if free < 24 << 30:
        pytest.skip("needs ~24 GB of device memory")
    rng = np.random.default_rng(9)
    n = 40_000
    t = pa.table({"big": rng.integers(-2**62, 2**62, n), "s": rng.integers(0, 7, n).astype(np.int16),
                  "v": pa.array(rng.integers(-2**40, 2**40, n), mask=rng.random(n) < 0.02)})
    st = util.make_storage({"t": t}, fragment_size=10_000)
    text = "SELECT big, s, COUNT(*), SUM(v), MIN(v) FROM t GROUP BY big, s"
    E = 150_000_000
    ex = Executor(st)
    ex.compact_threshold_bytes = 1 << 62
    pq = ex.plan(sql.parse(text, st.tables), E)
    assert pq.qmd.hash_type == abi.BASELINE_HASH and ex.lib.hdk_b200_buffer_size_bytes(C.byref(pq.qmd)) > (1 << 32)
    frags = st.get_table("t").fragments
    whole = ex.prepare(pq)
    ex.launch(pq, whole)
    torch.cuda.synchronize()
    assert int(whole["err"].item()) == 0
    parts = []
    for part in (frags[:2], frags[2:]):
        pp = ex.prepare(pq, fragments=part)
        ex.launch(pq, pp)
        torch.cuda.synchronize()
        assert int(pp["err"].item()) == 0
        parts.append(pp)
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    assert ex.lib.hdk_b200_reduce(C.byref(pq.plan), C.byref(pq.qmd), parts[0]["out"].data_ptr(), parts[1]["out"].data_ptr(), E,
                                  err.data_ptr(), None) == 0
    torch.cuda.synchronize()
    assert int(err.item()) == 0
    # entries beyond the 4 GiB offset exist in the whole-table buffer (keys are spread uniformly)
    row_bytes = whole["out"].numel() // E
    first_key = whole["out"].view(torch.int64)[:: row_bytes // 8]
    far = first_key[(1 << 32) // row_bytes + 1:]
    assert int((far != abi.EMPTY_KEY_64).sum().item()) > 100
    # decoded rows: whole == reduced halves == oracle over a small table of the same plan
    rows_whole = ResultSet.from_compact(pq, ex.compact_on_device(pq, whole["out"])).to_arrow().sort_by([("big", "ascending"), ("s", "ascending")])
    rows_red = ResultSet.from_compact(pq, ex.compact_on_device(pq, parts[0]["out"])).to_arrow().sort_by([("big", "ascending"), ("s", "ascending")])
    assert rows_whole.equals(rows_red) and rows_whole.num_rows > 30_000
    pq_small = util.plan_sql(st, text, max_groups_buffer_entry_count=131072)
    obuf, oerr = util.run_oracle(oracle_mod, st, pq_small)
    assert oerr == 0
    exp = util.sort_rows(util.result_columns(oracle_mod, pq_small, obuf), 2)
    got = util.arrow_rows(rows_whole)
    util.assert_rows_equal(got, exp)


@pytest.mark.parametrize("config", ["c1", "c3", "c5", "c4"])
def test_other_configs_at_baseline_size(config):
    """BASELINE.json configs[0], [2], [4], [3] at their full sizes on one GPU (TPC-H Q1 on 600 M lineitem rows, the star join
    on 2 B x 10 M rows, the composite-key group-by on 1 B rows / 100 M groups — through the radix-partitioned aggregation),
    checked through size-independent properties (benchcfg.check_*): conservation of rows, per-group results against
    independent torch reductions of the same columns (integers bit-exact, fp64 within 1e-9), checksum of keys x counts
    modulo 2^64, no key in two entries."""
    import torch
    import benchcfg
    free, _ = torch.cuda.mem_get_info()
    if config == "c4" and free < 90e9:
        pytest.skip("needs ~75 GB of device memory")
    out = benchcfg.per_config_single_gpu(torch.device("cuda", 0), 6537.0, reps=1, cpu=False, only=(config,))
    assert out, "nothing ran"
    for name, d in out.items():
        assert d["parity_check"].startswith("ok"), f"{name}: {d['parity_check']}"
        assert d["precompiled_shape"], name
    if config == "c4":
        from hdk_b200 import abi
        assert out["c4_baseline_hash"]["strategy"] == abi.STRATEGY_PARTITIONED
