"""The pre-compiled plan shapes (static_shapes.inc) on the benchmark's own generators at reduced size:
every named config must (a) dispatch to its pre-compiled kernel, (b) match the CPU oracle, (c) match the
generic (interpreting) kernel bit for bit on integer results."""
import numpy as np
import pytest

from hdk_b200 import abi
from tests import util

pytestmark = pytest.mark.gpu


def _oracle_frs(oracle, st, pq):
    return util.oracle_inputs(oracle, st, pq)


CASES = [
    ("taxi_q1", "taxi", "q1", 1), ("taxi_q2", "taxi", "q2", 1), ("taxi_q3", "taxi", "q3", 2), ("taxi_q4", "taxi", "q4", 3),
    ("c1_int64", "c1", "int", 1), ("c1_fp64", "c1", "fp", 1), ("tpch_q1", "lineitem", None, 2), ("star_join_sum", "star", None, 1),
    ("tpch_q6", "lineitem", "q6", 0), ("composite_key_sum_count", "c4", None, 2),
]


@pytest.mark.parametrize("name,kind,sub,nk", CASES)
def test_static_shape_parity(oracle_mod, name, kind, sub, nk):
    import torch
    import benchdata
    from hdk_b200 import sql
    from hdk_b200.executor import Executor
    from hdk_b200.storage import ArrowStorage
    dev = torch.device("cuda", 0)
    st = ArrowStorage()
    if kind == "taxi":
        benchdata.make_taxi(st, dev, 300_007, fragment_rows=70_001, keep_host=True)
        text = benchdata.TAXI_QUERIES[sub].split(" ORDER BY")[0]
    elif kind == "c1":
        benchdata.make_c1(st, dev, rows=250_003, fragment_rows=62_501, keep_host=True)
        text = benchdata.C1_QUERY if sub == "int" else benchdata.C1_QUERY_F
    elif kind == "lineitem":
        benchdata.make_lineitem(st, dev, 200_003, fragment_rows=50_001, keep_host=True)
        text = benchdata.TPCH_Q6 if sub == "q6" else benchdata.TPCH_Q1
    elif kind == "c4":
        benchdata.make_c4(st, dev, 200_003, 30_000, fragment_rows=50_001, keep_host=True)
        text = benchdata.C4_QUERY
    else:
        benchdata.make_star(st, dev, 300_007, 5_000, fragment_rows=70_001, keep_host=True)
        text = benchdata.C5_QUERY
    ex = Executor(st)
    unit = sql.parse(text, st.tables)
    pq = ex.plan(unit, 65536 if kind == "c4" else None)
    prep = ex.prepare(pq)
    info = ex.launch(pq, prep)
    torch.cuda.synchronize()
    assert int(prep["err"].item()) == 0
    assert (pq.qmd.hash_type == abi.BASELINE_HASH) == (kind == "c4")
    assert info.variant > 0, f"{name}: expected a pre-compiled kernel, ran the generic one"
    if name == "tpch_q1":
        assert info.strategy == 4, "TPC-H Q1 (6 groups, 11 accumulators) should keep its accumulators in registers"
    got = prep["out"].cpu().numpy().copy()
    # the same plan on the generic kernel
    from hdk_b200 import _lib
    _lib.debug_set("force_generic", 1)
    try:
        info2 = ex.launch(pq, prep)
    finally:
        _lib.debug_set("force_generic", 0)
    torch.cuda.synchronize()
    assert info2.variant == 0
    gen = prep["out"].cpu().numpy().copy()
    obuf, oerr = util.run_oracle(oracle_mod, st, pq, n_threads=4)
    assert oerr == 0
    exp = util.sort_rows(util.result_columns(oracle_mod, pq, obuf), nk)
    util.assert_rows_equal(util.sort_rows(util.result_columns(oracle_mod, pq, got), nk), exp)
    util.assert_rows_equal(util.sort_rows(util.result_columns(oracle_mod, pq, gen), nk), exp)
    has_fp_sum = any(ti.agg in (abi.AGG_SUM, abi.AGG_AVG) and ti.arg_type is not None and ti.arg_type.is_fp for ti in pq.infos)
    if not has_fp_sum and pq.qmd.hash_type == abi.PERFECT_HASH:   # (baseline hash: colliding keys may settle in either order)
        assert np.array_equal(got, obuf) and np.array_equal(gen, obuf)


@pytest.mark.parametrize("sub", ["int", "fp"])
def test_register_strategy_few_groups(oracle_mod, sub):
    """<= 8 groups and several wide aggregates: the REGISTER strategy (per-thread register accumulators) must match
    the oracle and the shared-memory strategies, NULL arguments, MIN/MAX and empty groups included."""
    import pyarrow as pa
    import torch
    from hdk_b200 import sql
    from hdk_b200.executor import Executor
    rng = np.random.default_rng(21)
    n = 150_011
    k = rng.choice([0, 1, 2, 4, 5], n, p=[0.5, 0.3, 0.15, 0.04, 0.01]).astype(np.int32)      # group 3 stays empty
    t = pa.table({"k": k,   # (no NULL keys: the pre-compiled c1 shapes have a NULL-free key)
                  "v": pa.array(rng.integers(-2**40, 2**40, n), mask=rng.random(n) < 0.02),
                  "f": pa.array(rng.uniform(-1e6, 1e6, n), mask=rng.random(n) < 0.02)})
    st = util.make_storage({"c1": t}, fragment_size=40_009)
    col = "v" if sub == "int" else "f"
    text = f"SELECT k, COUNT(*), SUM({col}), MIN({col}), MAX({col}) FROM c1 GROUP BY k"
    ex = Executor(st)
    pq = ex.plan(sql.parse(text, st.tables))
    prep = ex.prepare(pq)
    info = ex.launch(pq, prep)
    torch.cuda.synchronize()
    assert int(prep["err"].item()) == 0 and info.variant > 0 and info.strategy == 4
    got = prep["out"].cpu().numpy().copy()
    obuf, oerr = util.run_oracle(oracle_mod, st, pq, n_threads=2)
    assert oerr == 0
    exp = util.sort_rows(util.result_columns(oracle_mod, pq, obuf), 1)
    util.assert_rows_equal(util.sort_rows(util.result_columns(oracle_mod, pq, got), 1), exp)
    if sub == "int":
        assert np.array_equal(got, obuf)
    from hdk_b200 import _lib
    for strategy in (0, 1, 4):
        _lib.debug_set("force_strategy", strategy)
        try:
            i2 = ex.launch(pq, prep)
        finally:
            _lib.debug_set("force_strategy", -1)
        torch.cuda.synchronize()
        assert i2.strategy == strategy
        util.assert_rows_equal(util.sort_rows(util.result_columns(oracle_mod, pq, prep["out"].cpu().numpy()), 1), exp)


# (256, 512, 32768: partition counts that once gave a scatter level 256 destinations — the last one collided with the 0xff
#  "padding" mark of a tile's per-vector destination byte and its records were dropped)
@pytest.mark.parametrize("generic,slots,partitions", [(0, 0, 0), (1, 0, 0), (0, 1024, 2), (0, 0, 3000), (0, 0, 256), (0, 0, 512), (0, 0, 32768)])
def test_c4_partitioned_aggregation(oracle_mod, generic, slots, partitions):
    """Config 4's shape through the radix-partitioned aggregation (what runs at BASELINE.json's size, where the 2e8-entry
    table is far beyond L2): direct passes (plain columns, no interpreter) and the interpreting ones, against the oracle."""
    import torch
    import benchdata
    from hdk_b200 import _lib, sql
    from hdk_b200.executor import Executor
    from hdk_b200.storage import ArrowStorage
    dev = torch.device("cuda", 0)
    st = ArrowStorage()
    benchdata.make_c4(st, dev, 400_003, 60_000, fragment_rows=90_001, keep_host=True)
    knobs = (("partitioned_aggregation", 1, -1), ("force_generic", generic, 0), ("partitioned_table_slots", slots, 0),
             ("partitioned_partitions", partitions, 0))
    for k, v, _ in knobs:
        _lib.debug_set(k, v)
    try:
        ex = Executor(st)
        pq = ex.plan(sql.parse(benchdata.C4_QUERY, st.tables), 131072)
        prep = ex.prepare(pq)
        info = ex.launch(pq, prep)
        torch.cuda.synchronize()
    finally:
        for k, _, v in knobs:
            _lib.debug_set(k, v)
    assert int(prep["err"].item()) == 0
    assert info.strategy == abi.STRATEGY_PARTITIONED and (info.variant > 0) == (not generic)
    got = prep["out"].cpu().numpy()
    obuf, oerr = util.run_oracle(oracle_mod, st, pq, n_threads=4)
    assert oerr == 0
    util.assert_rows_equal(util.sort_rows(util.result_columns(oracle_mod, pq, got), 2), util.sort_rows(util.result_columns(oracle_mod, pq, obuf), 2))
