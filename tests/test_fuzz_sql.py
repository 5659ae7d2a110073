"""Differential fuzzing of the row evaluator, the reference's way (same SQL on SQLite, SQLiteComparator.cpp:66-170): random
aggregate queries — arithmetic over every integer width and fp32 / fp64, three-valued logic, IN / BETWEEN / IS NULL / CASE
(with and without ELSE), dictionary literals, expression group keys — over the reference's `test` fixture.  CPU: the oracle
(reference runtime) vs SQLite; GPU: the CUDA path through the façade vs SQLite.  Queries whose arithmetic overflows are
skipped (SQLite promotes to REAL, the reference raises ERR_OVERFLOW_OR_UNDERFLOW)."""
import random

import pytest

from tests import util

INT = ["x", "w", "y", "z", "t", "fx", "u", "smallint_nulls", "ofd"]
FP = ["f", "d", "dn", "ff", "fn"]


class Gen:
    def __init__(self, seed):
        self.r = random.Random(seed)

    def iexpr(self, d=0):
        r, p = self.r, self.r.random()
        if d > 2 or p < 0.35:
            return r.choice(INT)
        if p < 0.5:
            return f"({r.randint(-50, 50)})"
        if p < 0.8:
            return f"({self.iexpr(d + 1)} {r.choice('+-*')} {self.iexpr(d + 1)})"
        if p < 0.9:
            return f"(-({self.iexpr(d + 1)}))"
        if r.random() < 0.7:
            return f"CASE WHEN {self.bexpr(d + 1)} THEN {self.iexpr(d + 1)} ELSE {self.iexpr(d + 1)} END"
        return f"CASE WHEN {self.bexpr(d + 1)} THEN {self.iexpr(d + 1)} END"

    def fexpr(self, d=0):
        r, p = self.r, self.r.random()
        if d > 2 or p < 0.4:
            return r.choice(FP)
        if p < 0.5:
            return f"({r.uniform(-5, 5):.2f})"
        if p < 0.85:
            return f"({self.fexpr(d + 1)} {r.choice('+-*')} {r.choice([self.fexpr, self.iexpr])(d + 1)})"
        return f"CASE WHEN {self.bexpr(d + 1)} THEN {self.fexpr(d + 1)} ELSE {self.fexpr(d + 1)} END"

    def bexpr(self, d=0):
        r, p = self.r, self.r.random()
        if d > 2 or p < 0.45:
            e = r.choice([self.iexpr, self.fexpr])
            return f"{e(d + 1)} {r.choice(['<', '<=', '>', '>=', '=', '<>'])} {e(d + 1)}"
        if p < 0.55:
            return f"{r.choice(INT + FP)} IS {'NOT ' if r.random() < 0.5 else ''}NULL"
        if p < 0.65:
            vals = ", ".join(str(r.randint(-10, 110)) for _ in range(r.randint(1, 4)))
            return f"{r.choice(INT)} {'NOT ' if r.random() < 0.3 else ''}IN ({vals})"
        if p < 0.72:
            return f"{r.choice(INT)} BETWEEN {r.randint(-100, 50)} AND {r.randint(0, 1100)}"
        if p < 0.8:
            return f"str {r.choice(['=', '<>'])} '{r.choice(['foo', 'bar', 'baz', 'nope'])}'"
        if p < 0.9:
            return f"NOT ({self.bexpr(d + 1)})"
        return f"({self.bexpr(d + 1)} {r.choice(['AND', 'OR'])} {self.bexpr(d + 1)})"

    def query(self):
        r = self.r
        aggs = []
        for _ in range(r.randint(1, 3)):
            a = r.choice(["COUNT", "SUM", "MIN", "MAX", "AVG"])
            e = r.choice([self.iexpr, self.fexpr])() if a != "COUNT" or r.random() < 0.5 else "*"
            aggs.append(f"{a}({e})")
        keys = r.sample(["x", "z", "y", "w", "str", "fx", "smallint_nulls", f"CASE WHEN {self.bexpr(1)} THEN 1 ELSE 0 END", "(x + w)"],
                        r.randint(0, 2))
        q = f"SELECT {', '.join(keys + aggs)} FROM test"
        if r.random() < 0.7:
            q += f" WHERE {self.bexpr()}"
        if keys:
            q += " GROUP BY " + ", ".join(str(i + 1) for i in range(len(keys)))
        return q


def queries(seed, n):
    g = Gen(seed)
    return [g.query() for _ in range(n)]


@pytest.mark.parametrize("seed", [11, 12])
def test_fuzz_oracle_vs_sqlite(oracle_mod, seed):
    from hdk_b200 import planner
    from tests.test_sqlite_oracle import decode_with_dictionaries, reference_test_table
    tables = reference_test_table()
    st = util.make_storage(tables, fragment_size=3)
    compared = 0
    for text in queries(seed, 150):
        try:
            pq = util.plan_sql(st, text)
        except planner.UnsupportedPlan:
            continue
        buf, err = util.run_oracle(oracle_mod, st, pq, kind="reference")
        if err != 0:
            continue
        got = sorted(decode_with_dictionaries(st, pq, buf), key=repr)
        exp = sorted(util.sqlite_rows(tables, text, 0), key=repr)
        try:
            util.assert_rows_equal(got, exp, rel=1e-5)
        except AssertionError as e:
            raise AssertionError(f"{text}: {e}")
        compared += 1
    assert compared > 100


@pytest.mark.gpu
@pytest.mark.parametrize("seed,config", [(11, {}), (12, {}), (13, {}), (14, {"enable_columnar_output": True}),
                                         (15, {"baseline_threshold": 1}), (16, {"baseline_threshold": 1, "enable_columnar_output": True}),
                                         (17, {"bigint_count": True})])
def test_fuzz_gpu_vs_sqlite(seed, config):
    """(config: the planner options that pick the buffer layout — columnar output, baseline hash for every multi-key
    group-by, 64-bit COUNT — so that the generic kernel, finalize and the result decoding are fuzzed on each of them)"""
    import hdk_b200.hdk as hdk_mod
    from hdk_b200 import planner
    from hdk_b200.executor import QueryError
    from tests.test_sqlite_oracle import reference_test_table
    tables = reference_test_table()
    h = hdk_mod.init(**config)
    h.import_arrow(tables["test"], "test", fragment_size=3)
    compared = 0
    for text in queries(seed, 150):
        try:
            got = util.arrow_rows(h.sql(text).to_arrow())
        except planner.UnsupportedPlan:
            continue
        except QueryError:
            continue
        exp = util.sqlite_rows(tables, text, 0)
        try:
            util.assert_rows_equal(sorted(got, key=repr), sorted(exp, key=repr), rel=1e-5)
        except AssertionError as e:
            raise AssertionError(f"{text}: {e}")
        compared += 1
    assert compared > 100


# ---- joins: one-to-one and one-to-many perfect tables, composite keys (baseline tables, both layouts), two-join chains whose
# second key comes from the fact or from the first inner table, two one-to-many tables per plan; NULL-able keys on both sides
def join_tables(seed):
    import numpy as np
    import pyarrow as pa
    rng = np.random.default_rng(seed)
    n, m = 400, 60
    fact = pa.table({"a": rng.integers(0, 30, n).astype(np.int32),
                     "b": pa.array(rng.integers(0, 4, n).astype(np.int16), mask=rng.random(n) < 0.1),
                     "c": pa.array(rng.integers(-5, 5, n).astype(np.int64), mask=rng.random(n) < 0.05),
                     "v": rng.integers(-100, 100, n).astype(np.int64), "f": np.round(rng.normal(0, 10, n), 2)})
    d1 = pa.table({"a": pa.array(rng.permutation(80)[:m].astype(np.int32)), "b": rng.integers(0, 4, m).astype(np.int16),
                   "w": rng.integers(0, 9, m).astype(np.int32),
                   "g": pa.array(rng.integers(0, 3, m).astype(np.int8), mask=rng.random(m) < 0.1)})
    d2 = pa.table({"a": rng.integers(0, 35, m).astype(np.int32),
                   "b": pa.array(rng.integers(0, 5, m).astype(np.int16), mask=rng.random(m) < 0.1),
                   "h": rng.integers(0, 4, m).astype(np.int32), "x": np.round(rng.normal(5, 2, m), 1)})
    d3 = pa.table({"c": pa.array(np.arange(-6, 6).astype(np.int64)), "p": rng.integers(0, 3, 12).astype(np.int32)})
    return {"fact": fact, "d1": d1, "d2": d2, "d3": d3}


def join_queries(seed, n):
    r = random.Random(seed)

    def one():
        joins, cols, fcols = [], ["f0.a", "f0.b", "f0.c", "f0.v"], ["f0.f"]
        p = r.random()
        if p < 0.35:
            joins.append("JOIN d1 j1 ON f0.a = j1.a" if r.random() < 0.6 else "JOIN d1 j1 ON f0.a = j1.a AND f0.b = j1.b")
            cols += ["j1.w", "j1.g"]
        elif p < 0.7:
            joins.append("JOIN d2 j1 ON f0.a = j1.a" if r.random() < 0.5 else "JOIN d2 j1 ON f0.b = j1.b AND f0.a = j1.a")
            cols += ["j1.h"]
            fcols += ["j1.x"]
        elif p < 0.87:
            joins.append("JOIN d1 j1 ON f0.a = j1.a")
            joins.append("JOIN d3 j2 ON f0.c = j2.c" if r.random() < 0.5 else "JOIN d3 j2 ON j1.w - 4 = j2.c")
            cols += ["j1.w", "j1.g", "j2.p"]
        else:
            # two one-to-many tables in one plan (nested matching sets): both keyed by the fact table, or the second
            # chained behind an inner column of the first
            joins.append("JOIN d2 j1 ON f0.a = j1.a")
            joins.append("JOIN d2 j2 ON f0.b = j2.b" if r.random() < 0.5 else "JOIN d2 j2 ON j1.h = j2.b")
            cols += ["j1.h", "j2.h", "j2.a"]
            fcols += ["j2.x"]

        def ie(d=0):
            q = r.random()
            if d > 1 or q < 0.5:
                return r.choice(cols)
            if q < 0.6:
                return f"({r.randint(-9, 9)})"
            return f"({ie(d + 1)} {r.choice('+-*')} {ie(d + 1)})"

        def be(d=0):
            q = r.random()
            if d > 1 or q < 0.6:
                return f"{ie(1)} {r.choice(['<', '<=', '>', '=', '<>'])} {ie(1)}"
            if q < 0.7:
                return f"{r.choice(cols)} IS NOT NULL"
            if q < 0.8:
                return f"{r.choice(fcols)} > {r.uniform(-5, 8):.1f}"
            return f"({be(d + 1)} {r.choice(['AND', 'OR'])} {be(d + 1)})"

        aggs = []
        for _ in range(r.randint(1, 3)):
            a = r.choice(["COUNT", "SUM", "MIN", "MAX", "AVG"])
            e = "*" if a == "COUNT" and r.random() < 0.5 else (r.choice(fcols) if r.random() < 0.3 else ie())
            aggs.append(f"{a}({e})")
        keys = r.sample(cols, r.randint(0, 2))
        s = f"SELECT {', '.join(keys + aggs)} FROM fact f0 {' '.join(joins)}"
        if r.random() < 0.6:
            s += f" WHERE {be()}"
        if keys:
            s += " GROUP BY " + ", ".join(str(i + 1) for i in range(len(keys)))
        return s
    return [one() for _ in range(n)]


def test_fuzz_joins_oracle_vs_sqlite(oracle_mod):
    from hdk_b200 import planner
    from tests.test_sqlite_oracle import decode_with_dictionaries
    tables = join_tables(21)
    st = util.make_storage(tables, fragment_size=90)
    compared = 0
    for text in join_queries(21, 150):
        try:
            pq = util.plan_sql(st, text)
        except planner.UnsupportedPlan:
            continue
        buf, err = util.run_oracle(oracle_mod, st, pq, kind="reference")
        if err != 0:
            continue
        try:
            util.assert_rows_equal(sorted(decode_with_dictionaries(st, pq, buf), key=repr), sorted(util.sqlite_rows(tables, text, 0), key=repr),
                                   rel=1e-6)
        except AssertionError as e:
            raise AssertionError(f"{text}: {e}")
        compared += 1
    assert compared > 100


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [21, 22, 23])
def test_fuzz_joins_gpu_vs_sqlite(seed):
    import hdk_b200.hdk as hdk_mod
    from hdk_b200 import planner
    from hdk_b200.executor import QueryError
    tables = join_tables(seed)
    h = hdk_mod.init()
    for name, t in tables.items():
        h.import_arrow(t, name, fragment_size=90)
    compared = 0
    for text in join_queries(seed, 150):
        try:
            got = util.arrow_rows(h.sql(text).to_arrow())
        except (planner.UnsupportedPlan, QueryError):
            continue
        try:
            util.assert_rows_equal(sorted(got, key=repr), sorted(util.sqlite_rows(tables, text, 0), key=repr), rel=1e-6)
        except AssertionError as e:
            raise AssertionError(f"{text}: {e}")
        compared += 1
    assert compared > 100


# ---- larger tables: many ragged fragments, group counts from 3 to 60,000 (every accumulation strategy of the generic kernel:
# per-thread bins, per-CTA shared table, global work table, baseline hash), NULLs in keys and arguments
def large_table(seed, n=180_000):
    import numpy as np
    import pyarrow as pa
    rng = np.random.default_rng(seed)
    nul = lambda a, p: pa.array(a, mask=rng.random(n) < p)   # noqa: E731
    return {"big": pa.table({
        "k3": rng.integers(0, 3, n).astype(np.int8), "k40": nul(rng.integers(-20, 20, n).astype(np.int16), 0.02),
        "k900": rng.integers(100, 1000, n).astype(np.int32), "k60k": nul(rng.integers(0, 60000, n).astype(np.int32), 0.01),
        "wide": rng.integers(-(1 << 40), 1 << 40, n).astype(np.int64) // 1000 * 1000,
        "i": nul(rng.integers(-1000, 1000, n).astype(np.int32), 0.05), "j": rng.integers(0, 100, n).astype(np.int64),
        "f": nul(np.round(rng.normal(0, 100, n), 3), 0.03), "g": rng.random(n).astype(np.float32)})}


def large_queries(seed, n):
    r = random.Random(seed)
    out = []
    for _ in range(n):
        keys = r.choice([["k3"], ["k40"], ["k900"], ["k60k"], ["k3", "k40"], ["k40", "k900"], ["wide"], ["k900", "k3", "k40"], []])
        aggs = r.sample(["COUNT(*)", "COUNT(i)", "SUM(i)", "SUM(j)", "MIN(i)", "MAX(j)", "AVG(f)", "SUM(f)", "MIN(f)", "MAX(g)", "AVG(g)", "SUM(i * j)",
                         "SUM(CASE WHEN j > 50 THEN i ELSE 0 END)", "MIN(f + g)"], r.randint(1, 5))
        text = f"SELECT {', '.join(keys + aggs)} FROM big"
        if r.random() < 0.6:
            text += " WHERE " + r.choice(["i > 0", "j < 70 AND f IS NOT NULL", "k40 IN (1, 2, 3, 5, 8) OR g > 0.9", "f < 50.5 AND i <> 7", "k60k < 30000",
                                          "NOT (j BETWEEN 10 AND 20)", "k3 = 1 AND g <= 0.25"])
        if keys:
            text += " GROUP BY " + ", ".join(keys)
        out.append((text, len(keys)))
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("seed,config", [(41, {}), (42, {"enable_columnar_output": True}), (43, {"baseline_threshold": 1}),
                                         (44, {"baseline_threshold": 1, "partitioned": (0, 0)}),
                                         (45, {"baseline_threshold": 1, "partitioned": (512, 7)}),
                                         (46, {"baseline_threshold": 1, "enable_columnar_output": True, "partitioned": (0, 40), "generic": 1})])
def test_fuzz_large_tables_gpu_vs_oracle(oracle_mod, seed, config):
    """("partitioned": (table slots, partitions) — baseline-hash plans run through the radix-partitioned aggregation;
    "generic": its interpreter passes even where the plan is a direct one)"""
    from hdk_b200 import _lib
    config = dict(config)
    pa, generic = config.pop("partitioned", None), config.pop("generic", 0)
    if pa is not None:
        _lib.debug_set("partitioned_aggregation", 1)
        _lib.debug_set("partitioned_table_slots", pa[0])
        _lib.debug_set("partitioned_partitions", pa[1])
        _lib.debug_set("force_generic", generic)
    try:
        _fuzz_large_tables(oracle_mod, seed, config, pa is not None)
    finally:
        for k, v in (("partitioned_aggregation", -1), ("partitioned_table_slots", 0), ("partitioned_partitions", 0), ("force_generic", 0)):
            _lib.debug_set(k, v)


def _fuzz_large_tables(oracle_mod, seed, config, want_partitioned):
    import torch
    from hdk_b200 import abi, planner, sql
    from hdk_b200.executor import Executor
    from tests.test_gpu_parity import check_against_oracle
    tables = large_table(seed)
    st = util.make_storage(tables, fragment_size=23_456)
    strategies = set()
    for text, nk in large_queries(seed, 30):
        ex = Executor(st, planner.Config(**config))
        pq = ex.plan(sql.parse(text, st.tables), 262144)
        prep = ex.prepare(pq)
        info = ex.launch(pq, prep)
        torch.cuda.synchronize()
        assert int(prep["err"].item()) == 0, text
        strategies.add(int(info.strategy))
        try:
            check_against_oracle(oracle_mod, st, pq, prep["out"].cpu().numpy(), nk)
        except AssertionError as e:
            raise AssertionError(f"{text} [strategy {info.strategy}]: {e}")
    assert len(strategies) >= 3
    assert (abi.STRATEGY_PARTITIONED in strategies) == want_partitioned
