"""The SQL subset of the façade (hdk_b200/sql.py): what parses into which ExecutionUnit, and what is rejected loudly
(never silently approximated)."""
import numpy as np
import pyarrow as pa
import pytest

from hdk_b200 import abi, ir, planner, sql
from tests import util


@pytest.fixture(scope="module")
def st():
    n = 50
    t = pa.table({"a": np.arange(n, dtype=np.int32) % 5, "b": np.arange(n, dtype=np.int16) % 3, "x": np.arange(n, dtype=np.int64),
                  "f": np.linspace(0, 1, n), "ts": pa.array((np.arange(n) * 86400000).astype("datetime64[ms]")), "s": pa.array(["u", "v"] * (n // 2))})
    d = pa.table({"a": np.arange(5, dtype=np.int32), "b": np.zeros(5, dtype=np.int16), "w": np.arange(5, dtype=np.int64)})
    return util.make_storage({"t": t, "d": d}, fragment_size=20)


def test_select_aliases_group_by_order_limit(st):
    u = sql.parse("SELECT a AS k, EXTRACT(YEAR FROM ts) AS y, COUNT(*) AS n, AVG(f) FROM t WHERE x >= 3 AND f < 0.9 "
                  "GROUP BY k, y ORDER BY n DESC, k LIMIT 7", st.tables)
    assert u.table == "t" and u.target_names == ["k", "y", "n", "EXPR$3"] and u.limit == 7
    assert len(u.groupby_exprs) == 2 and isinstance(u.groupby_exprs[1], ir.ExtractYear if hasattr(ir, "ExtractYear") else object)
    assert u.order_by == [(2, True, True), (0, False, False)]    # Calcite default: NULLs high
    assert len(u.quals) >= 1


def test_join_conditions(st):
    u = sql.parse("SELECT d.w, SUM(t.x) FROM t JOIN d ON t.a = d.a GROUP BY d.w", st.tables)
    assert len(u.joins) == 1 and u.joins[0].inner_table == "d" and u.joins[0].inner_key_column == "a" and not u.joins[0].more_keys
    u = sql.parse("SELECT d.w, SUM(t.x) FROM t INNER JOIN d ON d.a = t.a AND t.b = d.b GROUP BY d.w", st.tables)
    assert u.joins[0].inner_key_columns == ["a", "b"] and [k.column for k in u.joins[0].outer_keys] == ["a", "b"]
    pq = util.plan_sql(st, "SELECT d.w, SUM(t.x) FROM t JOIN d ON t.a = d.a AND t.b = d.b GROUP BY d.w")
    assert pq.plan.joins[0].n_key_exprs == 2 and pq.plan.joins[0].key_width == 4
    pq = util.plan_sql(st, "SELECT d.w, SUM(t.x) FROM t JOIN d ON t.a = d.a GROUP BY d.w")
    assert pq.plan.joins[0].n_key_exprs == 0                                   # perfect table
    pq = util.plan_sql(st, "SELECT d.w, SUM(t.x) FROM t JOIN d ON t.a = d.a GROUP BY d.w", cfg=planner.Config(max_perfect_join_entries=2))
    assert pq.plan.joins[0].n_key_exprs == 1                                   # range too wide for the configured limit


def test_non_grouped_aggregate_plans_as_one_keyless_entry(st):
    pq = util.plan_sql(st, "SELECT COUNT(*), SUM(x), MIN(f) FROM t WHERE a = 2")
    assert pq.plan.n_keys == 0 and pq.qmd.entry_count == 1 and pq.qmd.keyless == 1 and pq.qmd.hash_type == abi.PERFECT_HASH


def test_in_lists_and_dictionary_literals(st):
    u = sql.parse("SELECT a, COUNT(*) FROM t WHERE b IN (0, 2) GROUP BY a", st.tables)
    q = u.quals[0]
    assert isinstance(q, ir.Logic) and q.op == "or" and all(isinstance(c, ir.Cmp) for c in q.args)
    u = sql.parse("SELECT a, COUNT(*) FROM t WHERE b NOT IN (0, 2) GROUP BY a", st.tables)
    assert u.quals[0].op == "not" and u.quals[0].args[0].op == "or"
    d = st.get_table("t").columns["s"].dictionary
    u = sql.parse("SELECT a, COUNT(*) FROM t WHERE s = 'v' GROUP BY a", st.tables)
    assert isinstance(u.quals[0].rhs, ir.Const) and u.quals[0].rhs.value == d.index("v")
    u = sql.parse("SELECT a, COUNT(*) FROM t WHERE 'zz' = s GROUP BY a", st.tables)       # not in the dictionary: an id no row carries
    assert u.quals[0].lhs.value == len(d)


def test_case_lowering_guards_only_what_can_raise(st):
    pq = util.plan_sql(st, "SELECT a, SUM(CASE WHEN b > 0 THEN x ELSE 0 END) FROM t GROUP BY a")
    nodes = [pq.plan.exprs[i] for i in range(pq.plan.n_exprs)]
    assert sum(n.op == abi.OP_CASE for n in nodes) == 1 and all(n.guard == 0 for n in nodes)   # pure arms: no guards
    pq = util.plan_sql(st, "SELECT a, SUM(CASE WHEN b <> 0 THEN x / b WHEN b = 0 THEN 1 ELSE x / (b - 1) END) FROM t GROUP BY a")
    nodes = [pq.plan.exprs[i] for i in range(pq.plan.n_exprs)]
    divs = [n for n in nodes if n.op == abi.OP_DIV]
    assert len(divs) == 2 and all(d.guard > 0 for d in divs) and divs[0].guard != divs[1].guard
    first = nodes[divs[0].guard - 1]
    assert first.op == abi.OP_NE                                            # THEN 1 runs where WHEN 1 is true
    assert nodes[divs[1].guard - 1].op == abi.OP_AND                        # ELSE runs where neither WHEN was
    case = [n for n in nodes if n.op == abi.OP_CASE]
    assert len(case) == 2 and case[1].a == divs[0].guard - 1 and nodes[case[1].ival] is case[0]   # nested from the last arm outwards
    pq = util.plan_sql(st, "SELECT a, COUNT(CASE WHEN b > 0 THEN 1 END) FROM t GROUP BY a")       # implicit ELSE NULL
    case = [pq.plan.exprs[i] for i in range(pq.plan.n_exprs) if pq.plan.exprs[i].op == abi.OP_CASE]
    assert case[0].type.nullable == 1 and pq.plan.exprs[case[0].ival].ival == abi.int_null(4)
    # the same division inside and outside an arm: two nodes, only one of them guarded
    pq = util.plan_sql(st, "SELECT a, SUM(CASE WHEN b <> 0 THEN x / b ELSE 0 END), SUM(x / b) FROM t GROUP BY a")
    divs = [pq.plan.exprs[i] for i in range(pq.plan.n_exprs) if pq.plan.exprs[i].op == abi.OP_DIV]
    assert sorted(d.guard > 0 for d in divs) == [False, True]


@pytest.mark.parametrize("text", [
    "SELECT a, COUNT(*) FROM t WHERE CASE WHEN b > 0 THEN s ELSE s END = 'u' GROUP BY a",   # CASE over dictionary values
    "SELECT a, SUM(CASE WHEN b THEN 1 ELSE 0 END) FROM t GROUP BY a",    # non-boolean WHEN
    "SELECT a, SUM(CASE WHEN b > 0 THEN NULL END) FROM t GROUP BY a",    # only NULL values
    "SELECT a, COUNT(*) FROM t WHERE s < 'v' GROUP BY a",                # ordering on dictionary ids is not string ordering
    "SELECT a, COUNT(*) FROM t WHERE a = 'u' GROUP BY a",                # string literal against a number
    "SELECT a, COUNT(*) FROM t WHERE a IN () GROUP BY a",                # empty IN list
    "SELECT a, COUNT(*) FROM t GROUP BY a HAVING COUNT(*) > 1",          # HAVING
    "SELECT a, COUNT(DISTINCT b) FROM t GROUP BY a",                    # count distinct
    "SELECT a FROM t",                                                  # projection without aggregation
    "SELECT a, x FROM t GROUP BY a",                                    # non-key, non-aggregate target
    "SELECT f, COUNT(*) FROM t GROUP BY f",                             # floating-point group key
    "SELECT d.w, COUNT(*) FROM t LEFT JOIN d ON t.a = d.a GROUP BY d.w",  # outer join
    "SELECT d.w, COUNT(*) FROM t JOIN d ON t.a < d.a GROUP BY d.w",     # non-equi join
    "SELECT a, SUM(x) OVER (PARTITION BY a) FROM t",                    # window function
    "SELECT 'lit' AS k, COUNT(*) FROM t GROUP BY k",                    # string literal as a group key
    "SELECT a, MIN(s) FROM t GROUP BY a",                               # MIN over a dictionary-encoded string (reference: throws)
])
def test_unsupported_sql_is_rejected(st, text):
    with pytest.raises((planner.UnsupportedPlan, SyntaxError, KeyError, ValueError)):
        util.plan_sql(st, text)


def test_select_aliases_are_not_visible_in_where(st):
    # `x` in WHERE is the table's column, not the select-list alias of `a`; in GROUP BY / ORDER BY the alias wins
    u = sql.parse("SELECT a AS x, COUNT(*) AS n FROM t WHERE x > 3 GROUP BY x ORDER BY x", st.tables)
    assert u.quals[0].lhs.column == "x" and u.groupby_exprs[0].column == "a" and u.order_by[0][0] == 0


def test_builder_api_join_agg_sort_units():
    """The pyhdk-shaped builder (python/pyhdk/hdk.py:1606-1992 agg / join / sort): units are built without touching a GPU."""
    import hdk_b200.hdk as hdk_mod
    h = hdk_mod.init()
    h.import_arrow(pa.table({"a": np.arange(10, dtype=np.int32), "b": np.arange(10, dtype=np.int64) % 3, "v": np.arange(10.0)}), "t")
    h.import_arrow(pa.table({"a": np.arange(5, dtype=np.int32), "b": np.arange(5, dtype=np.int64) % 3, "w": np.arange(5, dtype=np.int32)}), "d")
    t, d = h.scan("t"), h.scan("d")
    n = t.join(d, "a").agg(["w"], s="sum(v)", c="count").sort(("s", "desc", "first"), "w")
    assert n.unit.joins[0].inner_key_columns == ["a"] and n.unit.target_names == ["w", "s", "c"]
    assert n.unit.order_by == [(1, True, True), (0, False, False)]            # NULLs last unless asked otherwise (hdk.py:1687-1689)
    n2 = t.join(d, ["a", "b"]).agg("w", {"m": "min(v)"})
    assert n2.unit.joins[0].inner_key_columns == ["a", "b"]
    stats = lambda ti, c: [h.storage.get_table("t"), h.storage.get_table("d")][ti].col_stats(c)   # noqa: E731
    pq = planner.build_query(n2.unit, stats, 10, h.config)
    assert pq.plan.n_joins == 1 and pq.plan.joins[0].n_key_exprs == 2         # composite key → baseline join table
    with pytest.raises(planner.UnsupportedPlan):
        t.join(d, "a", how="left")


def test_join_key_typing_rules(st):
    """ir.join_key_for: what may be joined on stored values, and what is refused rather than compared wrongly."""
    days = ir.SqlType("date", 8, True, date_in_days=True)
    ts_s, ts_ms = ir.SqlType("timestamp", 8, True, unit=1), ir.SqlType("timestamp", 8, True, unit=1000)
    i32, dic = ir.SqlType("int", 4, True), ir.SqlType("dict", 4, True)
    col = lambda t, w=4: ir.ColumnRef(0, "c", t, w)   # noqa: E731
    k = ir.join_key_for(col(days), days)
    assert isinstance(k, ir.ColumnRef) and k.type.kind == "int" and not k.type.date_in_days       # probes with the stored days
    assert ir.join_key_for(col(i32), i32) == col(i32)
    assert ir.join_key_for(col(ts_s, 8), ts_s) == col(ts_s, 8)
    for outer, inner in ((col(days), i32), (col(i32), days), (col(dic), dic), (col(ts_s, 8), ts_ms), (col(i32), ts_s)):
        with pytest.raises(NotImplementedError):
            ir.join_key_for(outer, inner)


def test_null_div_by_zero_makes_dependents_nullable(st):
    u = sql.parse("SELECT a, SUM(x / b), COUNT(*), MIN(CASE WHEN x / b > 1 THEN f ELSE 0.5 END) FROM t WHERE NOT (x / b = 2) GROUP BY a", st.tables)
    e = ir.with_null_div_by_zero(u.target_exprs[1])
    assert e.arg.null_on_zero and e.arg.type.nullable and e.type.nullable
    assert not ir.with_null_div_by_zero(u.target_exprs[2]).type.nullable                           # COUNT(*) stays NOT NULL
    q = ir.with_null_div_by_zero(u.quals[0])
    assert q.type.nullable and q.args[0].type.nullable and q.args[0].lhs.null_on_zero
    c = ir.with_null_div_by_zero(u.target_exprs[3])
    assert c.arg.arms[0][0].type.nullable                                                          # the WHEN became nullable
    pq = util.plan_sql(st, "SELECT a, SUM(x / b) FROM t GROUP BY a", cfg=planner.Config(null_div_by_zero=True))
    div = [pq.plan.exprs[i] for i in range(pq.plan.n_exprs) if pq.plan.exprs[i].op == abi.OP_DIV]
    assert div and all(d.aux & 2 and d.type.nullable and d.guard == 0 for d in div)


def test_baseline_join_key_width_follows_the_inner_columns():
    """BaselineJoinHashTable::getKeyComponentWidth (JHT/BaselineJoinHashTable.cpp:502-509): 8 bytes per component as soon
    as one INNER key column is wider than 4 bytes — an int32 outer column probing an int64 inner column must not build
    a 4-byte table (2^32 + 5 would alias with 5)."""
    n = 40
    t = pa.table({"a": np.arange(n, dtype=np.int32) % 7, "b": np.arange(n, dtype=np.int32) % 3, "x": np.arange(n, dtype=np.int64)})
    d = pa.table({"w": np.array([0, 1, 2, 3, 2**32 + 5, 5, 6], dtype=np.int64), "b": np.arange(7, dtype=np.int32) % 3,
                  "g": np.arange(7, dtype=np.int32)})
    st2 = util.make_storage({"t": t, "d": d}, fragment_size=16)
    pq = util.plan_sql(st2, "SELECT d.g, COUNT(*) FROM t JOIN d ON t.a = d.w AND t.b = d.b GROUP BY d.g")
    assert pq.plan.joins[0].n_key_exprs == 2 and pq.plan.joins[0].key_width == 8
    # narrow on both sides stays at 4
    pq = util.plan_sql(st2, "SELECT d.g, COUNT(*) FROM t JOIN d ON t.a = d.g AND t.b = d.b GROUP BY d.g")
    assert pq.plan.joins[0].n_key_exprs == 2 and pq.plan.joins[0].key_width == 4


@pytest.mark.parametrize("text", [
    "SELECT d.w, COUNT(*) FROM t JOIN d ON t.f = d.a GROUP BY d.w",              # floating-point outer join key
    "SELECT d.w, COUNT(*) FROM t JOIN d ON CAST(t.x AS INT) = d.a GROUP BY d.w AND 1 = 1",
    "SELECT a, COUNT(*) FROM t GROUP BY a ORDER BY 0",                          # ORDER BY ordinal out of range
    "SELECT a, COUNT(*) FROM t GROUP BY a ORDER BY 3",
    "SELECT a, SUM(CASE ELSE 1 END) FROM t GROUP BY a",                         # CASE without WHEN
])
def test_refused_as_unsupported_plan(st, text):
    """Refusals must surface as UnsupportedPlan (the façade's "not on the hot path"), not as IndexError / a silently wrong
    answer: fp join keys (the reference refuses to hash-join on them), ORDER BY ordinals outside the select list."""
    with pytest.raises((planner.UnsupportedPlan, SyntaxError)):
        util.plan_sql(st, text)


def test_fp_join_key_is_unsupported_plan(st):
    with pytest.raises(planner.UnsupportedPlan):
        util.plan_sql(st, "SELECT d.w, COUNT(*) FROM t JOIN d ON t.f = d.a GROUP BY d.w")
    with pytest.raises(planner.UnsupportedPlan):
        util.plan_sql(st, "SELECT a, COUNT(*) FROM t GROUP BY a ORDER BY 0")
    with pytest.raises(planner.UnsupportedPlan):
        util.plan_sql(st, "SELECT a, COUNT(*) FROM t GROUP BY a ORDER BY 3")
