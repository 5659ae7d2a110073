"""ORDER BY / LIMIT over aggregated results.  The oracle's restatement of ResultSetComparator (QE/ResultSetSort.cpp:333-470)
is pinned against SQLite's ORDER BY here (CPU); tests/test_gpu_sort.py compares the device sort with it."""
import sqlite3
import struct

import numpy as np
import pytest

FLT_NULL = struct.unpack("<f", struct.pack("<I", 0x00800000))[0]
DBL_NULL = 2.2250738585072014e-308


def sort_case(n=500, seed=4):
    """Decoded result columns as hdk_b200_compact_result writes them: int64 / double cells with in-band NULLs."""
    rng = np.random.default_rng(seed)
    a = rng.integers(-5, 6, n).astype(np.int64)                      # many ties
    a[rng.random(n) < 0.1] = -(1 << 31)                               # int32 NULLs
    b = np.round(rng.normal(0, 3, n), 1)
    b[rng.random(n) < 0.15] = DBL_NULL
    b[rng.random(n) < 0.05] = -0.0
    c = rng.integers(-(1 << 62), 1 << 62, n).astype(np.int64)         # wide: every radix digit varies
    d = rng.integers(0, 4, n).astype(np.int64)                        # dictionary ids
    d[rng.random(n) < 0.1] = -(1 << 31)
    f = rng.integers(0, 3, n).astype(np.float64) * 0.5                # fp32-typed target
    f[rng.random(n) < 0.2] = FLT_NULL
    cols = [a, b, c, d, f]
    meta = [dict(is_fp=0, type_width=4), dict(is_fp=1, type_width=8), dict(is_fp=0, type_width=8),
            dict(is_fp=0, type_width=4, dictionary=["pear", "apple", "fig", "date"]), dict(is_fp=1, type_width=4)]
    return cols, meta


ORDERS = [
    [(0, False, False)], [(0, True, True)], [(0, False, True)], [(0, True, False)],
    [(1, False, False)], [(1, True, False)], [(1, True, True)],
    [(2, False, False)], [(2, True, False)],
    [(3, False, False)], [(3, True, True)],
    [(4, False, True)], [(4, True, False)],
    [(0, False, False), (1, True, True)], [(3, True, False), (0, False, True), (2, True, False)],
    [(4, False, False), (3, False, True), (1, False, False), (0, True, True)],
]


def order_entries(order, meta):
    return [dict(column=c, is_desc=int(d), nulls_first=int(nf), nullable=1, **meta[c]) for c, d, nf in order]


def key_rows(cols, meta, order, perm):
    """the ORDER BY tuple of every row in permutation order, NULL → None, dictionary id → string"""
    out = []
    for r in perm:
        row = []
        for c, _, _ in order:
            v = cols[c][r]
            m = meta[c]
            null = (FLT_NULL if m["type_width"] == 4 else DBL_NULL) if m["is_fp"] else -(1 << (8 * m["type_width"] - 1))
            if v == null:
                row.append(None)
            elif m.get("dictionary"):
                row.append(m["dictionary"][int(v)])
            else:
                row.append(float(v) + 0.0 if m["is_fp"] else int(v))
        out.append(tuple(row))
    return out


@pytest.mark.parametrize("order", ORDERS)
def test_comparator_restatement_vs_sqlite(oracle_mod, order):
    cols, meta = sort_case()
    n = len(cols[0])
    perm = oracle_mod.sort_permutation(cols, n, order_entries(order, meta))
    assert sorted(perm) == list(range(n))
    con = sqlite3.connect(":memory:")
    con.execute("CREATE TABLE t (c0, c1, c2, c3, c4)")
    con.executemany("INSERT INTO t VALUES (?,?,?,?,?)", key_rows(cols, meta, [(c, 0, 0) for c in range(5)], range(n)))
    clause = ", ".join(f"c{c} {'DESC' if d else 'ASC'} NULLS {'FIRST' if nf else 'LAST'}" for c, d, nf in order)
    exp = con.execute(f"SELECT {', '.join(f'c{c}' for c, _, _ in order)} FROM t ORDER BY {clause}").fetchall()
    assert key_rows(cols, meta, order, perm) == [tuple(r) for r in exp]
    top = oracle_mod.sort_permutation(cols, n, order_entries(order, meta), top_n=7)
    assert key_rows(cols, meta, order, top) == [tuple(r) for r in exp[:7]]
