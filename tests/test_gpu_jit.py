"""Run-time specialisation (csrc/jit.cu): plan shapes without pre-compiled kernels are instantiated by NVRTC from the same
scan_kernel.cuh and must give the interpreting kernel's — the oracle's — results.  "jit" = 2 compiles before the launch."""
import ctypes as C

import numpy as np
import pytest

from hdk_b200 import abi
from tests import util

pytestmark = pytest.mark.gpu


@pytest.fixture
def jit_sync():
    from hdk_b200 import _lib
    if not _lib.jit_stats()["available"]:
        pytest.skip("libnvrtc / libcu++ headers not found on this box")
    _lib.debug_set("jit", 2)
    yield _lib
    _lib.debug_set("jit", 0)


@pytest.mark.parametrize("idx", [1, 5, 6, 8, 9, 10, 12, 15, 16, 17])
def test_jit_kernels_match_oracle(oracle_mod, jit_sync, idx):
    """Shapes of the parity list that have no pre-compiled kernel: nullable keys, fp32 aggregates, filters with three-valued
    logic, columnar output, baseline hash (8- and 4-byte keys), joins through the slot-ordered payload."""
    import torch
    from hdk_b200 import sql
    from hdk_b200.executor import Executor
    from tests.test_gpu_parity import QUERIES, check_against_oracle, gen_tables
    tables = gen_tables()
    st = util.make_storage(tables, fragment_size={"t": 7001, "dim": 100000, "dim_many": 100000})
    text, nk, kw = QUERIES[idx]
    ex = Executor(st, kw.get("cfg"))
    pq = ex.plan(sql.parse(text, st.tables), kw.get("max_groups_buffer_entry_count"), kw.get("output_columnar"))
    prep = ex.prepare(pq)
    info = ex.launch(pq, prep)
    torch.cuda.synchronize()
    assert int(prep["err"].item()) == 0
    # (a query of the list may coincide with a pre-compiled shape: QUERIES[1] is config 1's)
    assert info.variant > 0, "expected specialised kernels (run-time compiled or pre-compiled), ran the interpreter"
    check_against_oracle(oracle_mod, st, pq, prep["out"].cpu().numpy(), nk)
    st_ = jit_sync.jit_stats()
    assert st_["shapes_failed"] == 0
    if info.variant == abi.VARIANT_JIT:
        assert st_["shapes_compiled"] >= 1 and st_["last_compile_ms"] > 0


def test_jit_in_the_background_then_specialised(oracle_mod):
    """Default mode: the first launch of a new shape runs the interpreting kernel while a worker thread compiles; once the
    shape is ready launches use it.  Same buffer both times (integer aggregates: byte-identical)."""
    import torch
    from hdk_b200 import _lib, sql
    from hdk_b200.executor import Executor
    from tests.test_gpu_parity import gen_tables
    if not _lib.jit_stats()["available"]:
        pytest.skip("libnvrtc not found")
    tables = gen_tables()
    st = util.make_storage(tables, fragment_size={"t": 7001, "dim": 100000, "dim_many": 100000})
    ex = Executor(st)
    pq = ex.plan(sql.parse("SELECT s, k, COUNT(*), SUM(v), MIN(w), MAX(v) FROM t WHERE w > 3 GROUP BY s, k", st.tables))
    prep = ex.prepare(pq)
    _lib.debug_set("jit", 1)
    try:
        info1 = ex.launch(pq, prep)
        torch.cuda.synchronize()
        first = prep["out"].cpu().numpy().copy()
        assert info1.variant == 0
        _lib.check(_lib.lib().hdk_b200_jit_wait(), "jit_wait")
        info2 = ex.launch(pq, prep)
        torch.cuda.synchronize()
        assert info2.variant == abi.VARIANT_JIT
        assert int(prep["err"].item()) == 0 and np.array_equal(first, prep["out"].cpu().numpy())
    finally:
        _lib.debug_set("jit", 0)
    obuf, oerr = util.run_oracle(oracle_mod, st, pq)
    assert oerr == 0 and np.array_equal(first, obuf)


def test_jit_fuzz_vs_sqlite(jit_sync):
    """A slice of the differential fuzz (random aggregate queries over the reference's `test` fixture) with every shape
    compiled at run time, against SQLite."""
    import hdk_b200.hdk as hdk_mod
    from hdk_b200 import planner
    from hdk_b200.executor import QueryError
    from tests.test_fuzz_sql import queries
    from tests.test_sqlite_oracle import reference_test_table
    tables = reference_test_table()
    h = hdk_mod.init()
    h.import_arrow(tables["test"], "test", fragment_size=3)
    compared = 0
    for text in queries(31, 24):
        try:
            got = util.arrow_rows(h.sql(text).to_arrow())
        except (planner.UnsupportedPlan, QueryError):
            continue
        exp = util.sqlite_rows(tables, text, 0)
        try:
            util.assert_rows_equal(sorted(got, key=repr), sorted(exp, key=repr), rel=1e-5)
        except AssertionError as e:
            raise AssertionError(f"{text}: {e}")
        compared += 1
    st_ = jit_sync.jit_stats()
    assert compared > 12 and st_["shapes_failed"] == 0 and st_["launches"] >= compared


def test_harvested_reference_queries_run_specialised(jit_sync):
    """The queries harvested from the reference's own Select.* tests (tests/test_sqlite_oracle.py) with every plan shape
    compiled at run time BEFORE its launch ("jit" = 2), rows vs SQLite: the specialised kernels, not the interpreter, give
    the reference's answers.  Every 5th query by default (a compile costs ~2 s of host time per new shape);
    HDK_B200_JIT_FULL=1 runs all of them (the log of such a run is profiles/r2_jit_harvested_full.log)."""
    import os
    import hdk_b200.hdk as hdk_mod
    from hdk_b200 import planner
    from tests.test_sqlite_oracle import REFERENCE_HARVESTED_QUERIES, harvested_tables
    tables = harvested_tables()
    h = hdk_mod.init()
    for name, t in tables.items():
        h.import_arrow(t, name, fragment_size=2)
    texts = [q for qs in REFERENCE_HARVESTED_QUERIES.values() for q in qs]
    stride = 1 if os.environ.get("HDK_B200_JIT_FULL") else 5
    before = jit_sync.jit_stats()
    specialised = interpreted = 0
    for text in texts[::stride]:
        res = h.sql(text)
        got = util.arrow_rows(res.to_arrow())
        exp = util.sqlite_rows(tables, text, 0)
        if "ORDER BY" not in text.upper():
            got, exp = sorted(got, key=repr), sorted(exp, key=repr)
        try:
            util.assert_rows_equal(got, exp, rel=1e-6)
        except AssertionError as e:
            raise AssertionError(f"{text}: {e}")
        info = res.launch_info
        if info is not None and info.variant > 0:
            specialised += 1
        else:
            interpreted += 1      # one-to-many joins and queries that read no column stay on the interpreter / launch nothing
    after = jit_sync.jit_stats()
    print(f"harvested queries: {specialised} specialised, {interpreted} interpreted, "
          f"{after['shapes_compiled'] - before['shapes_compiled']} shapes compiled in {after['total_compile_ms'] - before['total_compile_ms']:.0f} ms")
    assert after["shapes_failed"] == before["shapes_failed"]
    assert specialised >= 0.8 * (specialised + interpreted)


@pytest.mark.parametrize("text,nk", util.COMPOSITE_JOIN_QUERIES + [(util.NON_GROUPED_QUERIES[3], 0)])
def test_jit_kernels_probe_baseline_join_tables(oracle_mod, jit_sync, text, nk):
    """Run-time specialised kernels over BASELINE join tables (composite keys; a single key whose range is too wide for a
    perfect table): the table kind and its component nodes are structure of the plan shape.  (Found by running every harvested
    query specialised: the shape text once left the components out and the kernel probed the table as a perfect one.)"""
    import torch
    from hdk_b200 import sql
    from hdk_b200.executor import Executor
    from tests.test_gpu_parity import check_against_oracle
    st = util.make_storage(util.composite_join_tables(), fragment_size=1300)
    ex = Executor(st)
    pq = ex.plan(sql.parse(text, st.tables))
    prep = ex.prepare(pq)
    info = ex.launch(pq, prep)
    torch.cuda.synchronize()
    assert int(prep["err"].item()) == 0
    assert any(pq.plan.joins[j].n_key_exprs > 0 for j in range(pq.plan.n_joins))
    assert info.variant == abi.VARIANT_JIT
    check_against_oracle(oracle_mod, st, pq, prep["out"].cpu().numpy(), max(nk, 1))
