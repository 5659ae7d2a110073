"""hdk_b200_sort_permutation / hdk_b200_gather_rows against the restated ResultSetComparator (oracle.sort_permutation,
pinned against SQLite in tests/test_sort.py), on edge sizes, and — at sizes the Python comparator cannot reach —
through sortedness + permutation properties checked with independent torch reductions."""
import ctypes as C

import numpy as np
import pytest

from hdk_b200 import _lib, abi
from tests.test_sort import DBL_NULL, ORDERS, key_rows, order_entries, sort_case

pytestmark = pytest.mark.gpu


def device_sort(torch, cols, entries, n, limit=None, top_n=0):
    L = _lib.lib()
    dev = torch.device("cuda", 0)
    dcols = [torch.from_numpy(np.ascontiguousarray(c).view(np.int64)).to(dev) for c in cols]
    oes = (abi.OrderEntry * len(entries))()
    keep = []
    for e, oe in zip(oes, entries):
        e.column, e.is_fp, e.type_width, e.nullable = oe["column"], oe["is_fp"], oe["type_width"], oe["nullable"]
        e.is_desc, e.nulls_first = oe["is_desc"], oe["nulls_first"]
        if oe.get("dictionary"):
            d = oe["dictionary"]
            rank = np.empty(len(d), dtype=np.int32)
            rank[np.array(sorted(range(len(d)), key=d.__getitem__))] = np.arange(len(d), dtype=np.int32)
            keep.append(torch.from_numpy(rank).to(dev))
            e.dict_rank, e.dict_size = keep[-1].data_ptr(), len(d)
    perm = torch.full((max(n, 1),), -1, dtype=torch.int32, device=dev)
    sb = L.hdk_b200_sort_scratch_bytes(n)
    scratch = torch.empty(sb, dtype=torch.uint8, device=dev)
    ptrs = (C.c_void_p * abi.MAX_TARGETS)(*[c.data_ptr() for c in dcols])
    n_sorted = C.c_uint64(0)
    _lib.check(L.hdk_b200_sort_permutation(ptrs, oes, len(entries), n, top_n, perm.data_ptr(), C.byref(n_sorted), scratch.data_ptr(), sb,
                                           None), "sort")
    torch.cuda.synchronize()
    assert max(top_n, 0) <= n_sorted.value <= n and (top_n or n_sorted.value == n)
    if limit is None:
        return perm[:n_sorted.value], dcols
    n_out = min(limit, n)
    out = [torch.empty(n_out, dtype=torch.int64, device=dev) for _ in dcols]
    optrs = (C.c_void_p * len(dcols))(*[o.data_ptr() for o in out])
    _lib.check(L.hdk_b200_gather_rows(ptrs, optrs, len(dcols), perm.data_ptr(), n_out, None), "gather")
    torch.cuda.synchronize()
    return perm[:n], out


@pytest.mark.parametrize("order", ORDERS)
def test_device_sort_vs_reference_comparator(oracle_mod, order):
    import torch
    cols, meta = sort_case()
    n = len(cols[0])
    entries = order_entries(order, meta)
    perm, _ = device_sort(torch, cols, entries, n)
    perm = perm.cpu().numpy().astype(np.int64)
    assert sorted(perm.tolist()) == list(range(n))                                # a permutation of the rows
    exp = oracle_mod.sort_permutation(cols, n, entries)
    assert key_rows(cols, meta, order, perm) == key_rows(cols, meta, order, exp)  # same order up to ties
    # LIMIT: the gathered rows are the first rows of that order, whole rows
    _, out = device_sort(torch, cols, entries, n, limit=9)
    for c in range(len(cols)):
        np.testing.assert_array_equal(out[c].cpu().numpy(), np.ascontiguousarray(cols[c]).view(np.int64)[perm[:9]])


@pytest.mark.parametrize("n", [0, 1, 31, 2047, 2049, 3071, 3072, 3073, 3072 * 592 + 1, 3_000_017])
def test_device_sort_edge_sizes(n):
    import torch
    rng = np.random.default_rng(n)
    v = rng.integers(-1000, 1000, n).astype(np.int64)
    w = rng.integers(-(1 << 63), (1 << 63) - 1, n, dtype=np.int64)
    entries = [dict(column=0, is_fp=0, type_width=8, nullable=0, is_desc=1, nulls_first=0),
               dict(column=1, is_fp=0, type_width=8, nullable=0, is_desc=0, nulls_first=0)]
    perm, _ = device_sort(torch, [v, w], entries, n)
    perm = perm.cpu().numpy().astype(np.int64)
    exp = np.lexsort((w, -v))                       # last key is the primary one; (v, w) pairs are distinct w.h.p.
    np.testing.assert_array_equal(v[perm], v[exp])
    np.testing.assert_array_equal(w[perm], w[exp])


def test_device_sort_constant_keys_run_no_pass():
    """All keys equal: no digit varies, so no radix pass runs and the permutation is the identity."""
    import torch
    n = 100_000
    v = np.full(n, 42, dtype=np.int64)
    perm, _ = device_sort(torch, [v], [dict(column=0, is_fp=0, type_width=8, nullable=1, is_desc=0, nulls_first=1)], n)
    np.testing.assert_array_equal(perm.cpu().numpy(), np.arange(n, dtype=np.int32))


def test_device_sort_is_stable_across_order_entries():
    """Sorting by (a, b) must order the rows of equal a by b: the passes of the later entry may not disturb it."""
    import torch
    rng = np.random.default_rng(8)
    n = 1 << 20
    a = rng.integers(0, 7, n).astype(np.int64)
    b = rng.normal(size=n)
    b[rng.random(n) < 0.01] = DBL_NULL
    entries = [dict(column=0, is_fp=0, type_width=8, nullable=1, is_desc=0, nulls_first=0),
               dict(column=1, is_fp=1, type_width=8, nullable=1, is_desc=1, nulls_first=1)]
    perm, _ = device_sort(torch, [a, b], entries, n)
    perm = perm.cpu().numpy().astype(np.int64)
    assert np.array_equal(np.sort(perm), np.arange(n))
    sa, sb = a[perm], b[perm]
    assert np.all(np.diff(sa) >= 0)
    same = np.diff(sa) == 0
    isnull = sb == DBL_NULL
    # inside a group: NULLs first, then descending
    assert not np.any(same & ~isnull[:-1] & isnull[1:])
    both = same & ~isnull[:-1] & ~isnull[1:]
    assert np.all(sb[:-1][both] >= sb[1:][both])


def test_device_sort_large_sortedness_and_permutation():
    """1e8-group result size (config 4's output): checked on the device with torch (independent of the library)."""
    import torch
    dev = torch.device("cuda", 0)
    free, _ = torch.cuda.mem_get_info()
    n = min(100_000_000, int(free // 80))
    g = torch.Generator(device=dev)
    g.manual_seed(3)
    col = torch.randint(-(1 << 40), 1 << 40, (n,), dtype=torch.int64, device=dev, generator=g)
    L = _lib.lib()
    oe = (abi.OrderEntry * 1)()
    oe[0].column, oe[0].is_fp, oe[0].type_width, oe[0].nullable, oe[0].is_desc, oe[0].nulls_first = 0, 0, 8, 1, 1, 0
    perm = torch.empty(n, dtype=torch.int32, device=dev)
    sb = L.hdk_b200_sort_scratch_bytes(n)
    scratch = torch.empty(sb, dtype=torch.uint8, device=dev)
    ptrs = (C.c_void_p * abi.MAX_TARGETS)(col.data_ptr())
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    _lib.check(L.hdk_b200_sort_permutation(ptrs, oe, 1, n, 0, perm.data_ptr(), None, scratch.data_ptr(), sb, None), "sort")
    t1.record()
    torch.cuda.synchronize()
    print(f"sorted {n} rows in {t0.elapsed_time(t1):.2f} ms")
    # LIMIT 100 of the same order through the prefilter: same leading rows (the keys are distinct w.h.p., ties compare equal)
    top = torch.empty(n, dtype=torch.int32, device=dev)
    n_sorted = C.c_uint64(0)
    t0.record()
    _lib.check(L.hdk_b200_sort_permutation(ptrs, oe, 1, n, 100, top.data_ptr(), C.byref(n_sorted), scratch.data_ptr(), sb, None), "top")
    t1.record()
    torch.cuda.synchronize()
    print(f"top-100 of {n} rows in {t0.elapsed_time(t1):.2f} ms ({n_sorted.value} candidates)")
    assert 100 <= n_sorted.value < n // 8
    assert torch.equal(col[top[:100].long()], col[perm[:100].long()])
    del scratch
    s = col[perm.long()]
    assert bool((s[:-1] >= s[1:]).all())                                           # descending
    assert int(s.sum().item()) == int(col.sum().item())                            # same multiset (checksum) …
    seen = torch.zeros(n, dtype=torch.uint8, device=dev)
    seen[perm.long()] = 1
    assert int(seen.sum().item()) == n                                             # … and every row exactly once


def test_order_by_limit_through_the_facade_vs_sqlite(oracle_mod):
    """hdk.sql with ORDER BY … [NULLS FIRST|LAST] LIMIT: compact → sort → gather on the device; rows vs SQLite in order."""
    import pyarrow as pa
    import hdk_b200.hdk as hdk_mod
    from tests import util
    rng = np.random.default_rng(12)
    n = 20_000
    t = pa.table({"k": pa.array(rng.integers(0, 300, n).astype(np.int32), mask=rng.random(n) < 0.02),
                  "g": rng.integers(0, 5, n).astype(np.int16),
                  "v": pa.array(rng.integers(-50, 50, n).astype(np.int64), mask=rng.random(n) < 0.3),
                  "s": pa.array(rng.choice(["kiwi", "apple", "fig", "banana"], n))})
    hdk = hdk_mod.init()
    hdk.import_arrow(t, "t", fragment_size=3000)
    queries = [
        "SELECT k, COUNT(*) AS n, SUM(v) AS sv FROM t GROUP BY k ORDER BY n DESC, k ASC NULLS FIRST LIMIT 17",
        "SELECT k, MIN(v) AS m FROM t GROUP BY k ORDER BY m ASC NULLS LAST, k DESC NULLS LAST",
        "SELECT s, g, AVG(v) AS a FROM t GROUP BY s, g ORDER BY s DESC, a ASC NULLS FIRST",
        "SELECT g, COUNT(v) AS c FROM t GROUP BY g ORDER BY g DESC LIMIT 3",
    ]
    for q in queries:
        res = hdk.sql(q)
        assert res.result_set.sorted_on_device
        got = util.arrow_rows(res.to_arrow())
        exp = util.sqlite_rows({"t": t}, q, 0)
        util.assert_rows_equal(got, exp, rel=1e-9)


@pytest.mark.parametrize("order", [ORDERS[1], ORDERS[5], ORDERS[9], ORDERS[13], ORDERS[14]])
@pytest.mark.parametrize("top_n", [1, 10, 5000])
def test_limit_prefilter_keeps_every_row_of_the_answer(oracle_mod, order, top_n):
    """LIMIT through the radix-select prefilter (n >= 65536, top_n <= n / 8): the first top_n ORDER BY tuples equal the
    full sort's, with heavy ties on the primary target (|values| ~ 11, NULLs) so the threshold bucket holds many rows."""
    import torch
    cols, meta = sort_case(n=100_000, seed=21)
    n = len(cols[0])
    entries = order_entries(order, meta)
    full, _ = device_sort(torch, cols, entries, n)
    part, _ = device_sort(torch, cols, entries, n, top_n=top_n)
    full, part = full.cpu().numpy().astype(np.int64), part.cpu().numpy().astype(np.int64)
    assert len(np.unique(part)) == len(part)
    assert key_rows(cols, meta, order, part[:top_n]) == key_rows(cols, meta, order, full[:top_n])


@pytest.mark.parametrize("seed", [5, 6])
def test_fuzz_order_by_limit_vs_sqlite(seed):
    """Random ORDER BY lists (1-3 targets: group keys incl. dictionary strings and NULL-able keys, COUNT / SUM / MIN / AVG
    results; ASC / DESC; NULLS FIRST / LAST) with and without LIMIT over the reference's `test` and `logical_size_test` fixtures:
    the sequence of ORDER BY tuples must equal SQLite's (rows that tie on every ORDER BY target may come in any order)."""
    import random
    import hdk_b200.hdk as hdk_mod
    from tests import util
    from tests.test_sqlite_oracle import harvested_tables
    tables = harvested_tables()
    h = hdk_mod.init()
    for name in ("test", "logical_size_test"):
        h.import_arrow(tables[name], name, fragment_size=4)
    r = random.Random(seed)
    shapes = [("test", ["x", "z", "str", "fx", "smallint_nulls", "w", "ss"], ["COUNT(*)", "SUM(y)", "MIN(fn)", "AVG(dn)", "MAX(t)", "SUM(u)"]),
              ("logical_size_test", ["id", "id_null", "small_int_null", "tiny_int", "big_int_null"],
               ["COUNT(*)", "SUM(float_null)", "AVG(big_int_null)", "MIN(tiny_int_null)", "MAX(double_null)"])]
    compared = 0
    for _ in range(60):
        table, keys, aggs = r.choice(shapes)
        ks = r.sample(keys, r.randint(1, 2))
        ags = r.sample(aggs, r.randint(1, 3))
        names = [f"k{i}" for i in range(len(ks))] + [f"a{i}" for i in range(len(ags))]
        sel = ", ".join(f"{e} AS {n}" for e, n in zip(ks + ags, names))
        order = r.sample(range(len(names)), r.randint(1, min(3, len(names))))
        clause = ", ".join(f"{names[i]} {r.choice(['ASC', 'DESC'])} NULLS {r.choice(['FIRST', 'LAST'])}" for i in order)
        text = f"SELECT {sel} FROM {table} GROUP BY {', '.join(names[:len(ks)])} ORDER BY {clause}"
        if r.random() < 0.5:
            text += f" LIMIT {r.randint(1, 6)}"
        res = h.sql(text)
        assert res.result_set.sorted_on_device
        got = [tuple(row.values()) for row in res.to_arrow().to_pylist()]
        exp = util.sqlite_rows(tables, text, 0)
        try:
            util.assert_rows_equal([tuple(g[i] for i in order) for g in got], [tuple(e[i] for i in order) for e in exp], rel=1e-6)
            if "LIMIT" not in text:
                util.assert_rows_equal(sorted(got, key=repr), sorted(exp, key=repr), rel=1e-6)
        except AssertionError as e:
            raise AssertionError(f"{text}: {e}")
        compared += 1
    assert compared == 60
