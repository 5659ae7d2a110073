"""Shared helpers of the test-suite: build tables, plan queries, run the CPU oracle, compare."""
import ctypes as C
import sqlite3

import numpy as np
import pyarrow as pa

from hdk_b200 import abi, planner, sql, storage


def make_storage(tables, fragment_size=1000):
    st = storage.ArrowStorage()
    for name, t in tables.items():
        fs = fragment_size[name] if isinstance(fragment_size, dict) else fragment_size
        st.import_arrow_table(t, name, fragment_size=fs)
    return st


def plan_sql(st, text, cfg=None, **kw):
    unit = sql.parse(text, st.tables)
    tabs = [st.get_table(unit.table)] + [st.get_table(j.inner_table) for j in unit.joins]
    return planner.build_query(unit, lambda ti, c: tabs[ti].col_stats(c), tabs[0].num_rows, cfg or planner.Config(), **kw)


_KEEP = []   # ctypes temporaries referenced by JoinColumn structs


def oracle_inputs(oracle, st, pq):
    """Fragments + host-built join tables / inner columns for the oracle (and query_host)."""
    outer = st.get_table(pq.unit.table)
    frs = oracle.Fragments([[fr.chunks[c] for c in pq.columns] for fr in outer.fragments], [fr.num_rows for fr in outer.fragments])
    join_tables, inner_cols = [], []
    for j, js in enumerate(pq.unit.joins):
        inner = st.get_table(js.inner_table)
        pj = pq.plan.joins[j]
        if pj.n_key_exprs > 0:
            # baseline join table built by the oracle's own builder (one-to-one layout, 2 x rows entries)
            kc, kw = pj.n_key_exprs, pj.key_width
            key_cols = js.inner_key_columns[:kc]
            rows = sum(f.num_rows for f in inner.fragments)
            E = 2 * max(rows, 1)
            jcs = (abi.JoinColumn * kc)()
            tis = (abi.JoinColumnTypeInfo * kc)()
            keep = []
            for k, kcname in enumerate(key_cols):
                ci = inner.columns[kcname]
                lo, hi, _ = inner.join_key_range(kcname)
                jc = oracle.make_join_column([f.chunks[kcname] for f in inner.fragments], ci.phys_width)
                keep.append(jc)
                jcs[k] = jc
                tis[k] = oracle.make_type_info(ci.phys_width, lo, hi, abi.int_null(ci.phys_width))
            L = oracle.lib()
            buf = np.empty(E * (kc + 1) * kw, dtype=np.uint8)
            L.oracle_init_baseline_hash_join_buff(buf.ctypes.data, E, kc, 1, -1, kw)
            rc = L.oracle_fill_baseline_hash_join_buff(buf.ctypes.data, E, -1, 0, kc, 1, jcs, tis, kw)
            pj.one_to_many = 0
            if rc == -1:
                # duplicate composite keys: dictionary without payload, then offsets | counts | payload behind it
                # (BaselineJoinHashTable one-to-many layout)
                dict_bytes = E * kc * kw
                buf = np.empty(dict_bytes + (2 * E + rows) * 4, dtype=np.uint8)
                L.oracle_init_baseline_hash_join_buff(buf.ctypes.data, E, kc, 0, -1, kw)
                assert L.oracle_fill_baseline_hash_join_buff(buf.ctypes.data, E, -1, 0, kc, 0, jcs, tis, kw) == 0
                assert L.oracle_fill_one_to_many_baseline_hash_table(buf.ctypes.data + dict_bytes, buf.ctypes.data, E, -1, kc,
                                                                     jcs, tis, kw) == 0
                pj.one_to_many = 1
            else:
                assert rc == 0
            pj.entry_count = E
            join_tables.append(buf)
            inner_cols.append([np.concatenate([f.chunks[c] for f in inner.fragments]) for c in pq.inner_columns[j]])
            _KEEP.append(keep)
            continue
        lo, hi, _ = inner.join_key_range(js.inner_key_column)
        ci = inner.columns[js.inner_key_column]
        E = hi - lo + 1
        chunks = [f.chunks[js.inner_key_column] for f in inner.fragments]
        jc = oracle.make_join_column(chunks, ci.phys_width)
        ti = oracle.make_type_info(ci.phys_width, lo, hi, abi.int_null(ci.phys_width))
        L = oracle.lib()
        buf = np.empty(E, dtype=np.int32)
        L.oracle_init_hash_join_buff(buf.ctypes.data, E, -1)
        rc = L.oracle_fill_hash_join_buff(buf.ctypes.data, -1, 0, C.byref(jc), C.byref(ti), 1)
        one_to_many = 0
        if rc == -1:    # NeedsOneToManyHash: offsets | counts | payload
            rows = sum(f.num_rows for f in inner.fragments)
            buf = np.empty(2 * E + rows, dtype=np.int32)
            L.oracle_fill_one_to_many_hash_table(buf.ctypes.data, E, -1, C.byref(jc), C.byref(ti), 1)
            one_to_many = 1
        else:
            assert rc == 0
        pj = pq.plan.joins[j]
        pj.one_to_many, pj.payload_by_slot, pj.min_key, pj.max_key, pj.entry_count = one_to_many, 0, lo, hi, E
        join_tables.append(buf)
        inner_cols.append([np.concatenate([f.chunks[c] for f in inner.fragments]) for c in pq.inner_columns[j]])
    return frs, join_tables, inner_cols


def run_oracle(oracle, st, pq, kind=None, n_threads=2, per_fragment=True):
    """The oracle over its own host-built join tables.  kind=None: the reference's own runtime (oracle/_ref, compiled from
    the reference's sources) when it is present — it travels to the GPU box — else the port pinned to it on the CPU.  oracle_inputs describes those tables in the plan's join PODs
    (layout, entry count, no slot-ordered payload), so it works on a copy: `pq` may already be prepared for the GPU,
    whose join tables are laid out differently."""
    import copy
    pq2 = copy.copy(pq)
    pq2.plan = abi.Plan.from_buffer_copy(pq.plan)
    if kind is None:
        kind = "reference" if oracle.ref_available() else "port"
    frs, jt, ic = oracle_inputs(oracle, st, pq2)
    buf, err = oracle.run_query(pq2, frs, jt, ic, n_threads=n_threads, kind=kind, per_fragment=per_fragment)
    return buf, err


def result_columns(oracle, pq, buf):
    vals, nulls = oracle.iterate(pq, buf)
    return oracle.rows_to_columns(pq, vals, nulls)


def sort_rows(cols, n_keys):
    """rows sorted by the first n_keys columns (NULLs first) → list of tuples with None for NULL"""
    n = len(cols[0]) if cols else 0
    rows = []
    for i in range(n):
        row = []
        for c in cols:
            m = np.ma.getmaskarray(c)[i]
            v = np.ma.getdata(c)[i]
            row.append(None if m else (float(v) if np.issubdtype(np.asarray(v).dtype, np.floating) else int(v)))
        rows.append(tuple(row))
    keyf = lambda r: tuple((0, 0) if x is None else (1, x) for x in r[:n_keys])  # noqa: E731
    return sorted(rows, key=keyf)


def assert_rows_equal(got, exp, rel=1e-9, float_cols=()):
    assert len(got) == len(exp), f"row count {len(got)} != {len(exp)}"
    for r, (g, e) in enumerate(zip(got, exp)):
        assert len(g) == len(e)
        for c, (a, b) in enumerate(zip(g, e)):
            if a is None or b is None:
                assert a is None and b is None, f"row {r} col {c}: {a} vs {b}"
            elif isinstance(a, float) or isinstance(b, float) or c in float_cols:
                assert abs(float(a) - float(b)) <= rel * max(abs(float(b)), 1e-300) + 1e-300, f"row {r} col {c}: {a} vs {b}"
            else:
                assert a == b, f"row {r} col {c}: {a} vs {b}"


def arrow_rows(table):
    """Rows of an Arrow table as tuples, column by column (Table.to_pylist() makes dicts: two result columns of the same
    name — `SELECT j1.h, j2.h …` — would collapse into one)."""
    cols = [c.to_pylist() for c in table.columns]
    return [tuple(r) for r in zip(*cols)] if cols else []


def sqlite_rows(tables, text, n_keys):
    """The reference's own differential oracle: the same SQL on SQLite
    (omniscidb/Tests/ArrowSQLRunner/SQLiteComparator.cpp:66-170)."""
    con = sqlite3.connect(":memory:")
    for name, t in tables.items():
        cols = t.column_names
        con.execute(f"CREATE TABLE {name} ({', '.join(cols)})")
        # temporal columns go in as the text the reference's own SQLite side holds ('2014-12-13 22:23:15[.fff]', '1999-09-09')
        data = list(zip(*[(t.column(c).cast(pa.string()) if pa.types.is_temporal(t.column(c).type) else t.column(c)).to_pylist()
                          for c in cols]))
        con.executemany(f"INSERT INTO {name} VALUES ({', '.join('?' * len(cols))})", data)
    rows = [tuple(r) for r in con.execute(text).fetchall()]
    keyf = lambda r: tuple((0, 0) if x is None else (1, x) for x in r[:n_keys])  # noqa: E731
    return sorted(rows, key=keyf)


def composite_join_tables(seed=1, n=5000):
    """Fact table + a dimension keyed by (a, b) with unique pairs + a dimension keyed by a 2^40-spaced int64
    (range too wide for a perfect table): inputs of the baseline-join tests."""
    rng = np.random.default_rng(seed)
    t = pa.table({"a": rng.integers(0, 40, n).astype(np.int32),
                  "b": pa.array(rng.integers(0, 30, n).astype(np.int16), mask=rng.random(n) < 0.02),
                  "x": rng.integers(-100, 100, n), "f": rng.integers(-10**6, 10**6, n) / 128.0,
                  "big": pa.array(rng.integers(0, 300, n) * (2**40), mask=rng.random(n) < 0.01)})
    m = 700
    pairs = rng.permutation(40 * 30)[:m]
    dim = pa.table({"a": (pairs // 30).astype(np.int32), "b": (pairs % 30).astype(np.int16),
                    "attr": rng.integers(0, 9, m).astype(np.int32), "w": rng.integers(0, 1000, m)})
    dim2 = pa.table({"big": rng.permutation(300)[:200] * (2**40), "g": rng.integers(0, 5, 200).astype(np.int32),
                     "u": rng.integers(0, 50, 200) / 8.0})
    return {"t": t, "dim": dim, "dim2": dim2}


COMPOSITE_JOIN_QUERIES = [
    ("SELECT dim.attr, COUNT(*), SUM(t.x), SUM(dim.w) FROM t JOIN dim ON t.a = dim.a AND t.b = dim.b GROUP BY dim.attr", 1),
    ("SELECT dim.attr, t.a, MIN(t.f), MAX(dim.w), AVG(t.x) FROM t JOIN dim ON t.a = dim.a AND t.b = dim.b WHERE dim.w > 100 GROUP BY dim.attr, t.a", 2),
    ("SELECT dim2.g, COUNT(*), SUM(t.f * dim2.u), SUM(t.x) FROM t JOIN dim2 ON t.big = dim2.big GROUP BY dim2.g", 1),
]


NON_GROUPED_QUERIES = [
    "SELECT COUNT(*), SUM(x), AVG(f), MIN(x), COUNT(x) FROM t WHERE a < 20",
    "SELECT SUM(x), MAX(f), COUNT(*) FROM t WHERE a > 1000",                     # no row passes: NULL, NULL, 0
    "SELECT SUM(f * (1 - f)), MIN(f), MAX(x) FROM t",
    "SELECT SUM(t.x), COUNT(*) FROM t JOIN dim ON t.a = dim.a AND t.b = dim.b WHERE dim.w > 500",
]


def wide_inner_key_tables(seed=5, n=3000):
    """An int32 outer key joined to an int64 inner column holding values >= 2^32 whose low halves collide with real outer
    values (and one whose low half is EMPTY_KEY_32): a 4-byte join table would alias them."""
    rng = np.random.default_rng(seed)
    t = pa.table({"a": rng.integers(0, 50, n).astype(np.int32), "b": rng.integers(0, 4, n).astype(np.int32), "x": rng.integers(-100, 100, n)})
    w = np.concatenate([np.arange(0, 30, dtype=np.int64), 2**32 + np.arange(30, 50, dtype=np.int64), [2**32 + 2**31 - 1, 2**33 + 7]])
    b = np.arange(len(w), dtype=np.int32) % 4
    dim = pa.table({"w": w, "b": b, "g": (np.arange(len(w)) % 6).astype(np.int32)})
    return {"t": t, "dim": dim}


WIDE_INNER_KEY_QUERY = ("SELECT dim.g, COUNT(*), SUM(t.x) FROM t JOIN dim ON t.a = dim.w AND t.b = dim.b GROUP BY dim.g", 1)


def layout_tables(seed=9, n=6000):
    """Inputs of the layout tests: an int key (keyless + columnar layouts), a DATE key (bucketed perfect hash: one bin per
    day) with NULLs, NOT NULL and nullable int64 arguments."""
    import datetime
    rng = np.random.default_rng(seed)
    d0 = (datetime.date(2021, 3, 1) - datetime.date(1970, 1, 1)).days
    t = pa.table({"k": rng.integers(10, 70, n).astype(np.int32),
                  "d": pa.array((d0 + rng.integers(0, 200, n)).astype("datetime64[D]"), mask=rng.random(n) < 0.03),
                  "pos": rng.integers(1, 5000, n),
                  "v": pa.array(rng.integers(-1000, 1000, n), mask=rng.random(n) < 0.1),
                  "f": rng.normal(0, 5, n)})
    fields = [pa.field(f.name, f.type, nullable=(f.name != "pos")) for f in t.schema]
    return {"t": pa.table([t.column(f.name) for f in fields], schema=pa.schema(fields))}


LAYOUT_QUERIES = [
    # (text, n_keys, columnar, compare with SQLite?)
    ("SELECT k, COUNT(*), SUM(v), MIN(f) FROM t GROUP BY k", 1, True, True),                     # keyless + columnar
    ("SELECT k, AVG(f), MAX(pos) FROM t WHERE v > -500 GROUP BY k", 1, True, True),              # keyless on AVG's count slot
    ("SELECT MIN(pos), COUNT(*), SUM(pos) FROM t GROUP BY k", 0, True, False),                   # the key lands in the MIN slot (reference quirk)
    ("SELECT d, COUNT(*), SUM(pos), MIN(v) FROM t GROUP BY d", 1, False, True),                  # DATE key: bucket = one day
    ("SELECT d, k, COUNT(*), AVG(f) FROM t GROUP BY d, k", 2, False, True),                      # bucketed key inside a multi-key perfect hash
    ("SELECT d, COUNT(*), MAX(f) FROM t WHERE k < 40 GROUP BY d", 1, True, True),                # bucketed + columnar
]
