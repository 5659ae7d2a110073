"""Pins the CPU oracle against the reference's own known-answer vectors and, when
oracle/_ref was built from the reference's sources, against the reference runtime itself.

Golden vectors restated here come from the reference's tests:
  omniscidb/Tests/JoinHashTableTest.cpp:133-267 (perfect 1:1 / 1:N), :355-442 (keyed), :444-650 (multi-fragment)
  omniscidb/QueryEngine/GroupByHashTest.cpp:57-265 (get_group_value semantics)
  omniscidb/Tests/PartitionedGroupByTest.cpp:39-139 (20-row multi-type-key sums, 1e12 key ⇒ baseline hash)
  python/tests/test_pyhdk_api.py:457-497 (agg), :609-672 (join)
"""
import ctypes as C

import numpy as np
import pyarrow as pa
import pytest

from hdk_b200 import abi
from tests import util

KINDS = ["port", "reference"]


def _lib(oracle, kind):
    if kind == "reference" and not oracle.ref_available():
        pytest.skip("oracle/_ref not built (reference tree absent)")
    return oracle.lib(kind)


# ------------------------------------------------------------------------- join hash tables ---
def _perfect(oracle, vals, frag=None):
    vals = np.array(vals, dtype=np.int32)
    chunks = [vals] if frag is None else [vals[i:i + frag] for i in range(0, len(vals), frag)]
    lo, hi = int(vals.min()), int(vals.max())
    E = hi - lo + 1
    jc = oracle.make_join_column(chunks, 4)
    ti = oracle.make_type_info(4, lo, hi, abi.int_null(4))
    L = oracle.lib()
    buf = np.empty(E, dtype=np.int32)
    L.oracle_init_hash_join_buff(buf.ctypes.data, E, -1)
    rc = L.oracle_fill_hash_join_buff(buf.ctypes.data, -1, 0, C.byref(jc), C.byref(ti), 1)
    if rc == 0:
        return "OneToOne", buf, E
    buf = np.empty(2 * E + len(vals), dtype=np.int32)
    L.oracle_fill_one_to_many_hash_table(buf.ctypes.data, E, -1, C.byref(jc), C.byref(ti), 1)
    return "OneToMany", buf, E


def decode_one_to_many(buf, E, lo=0):
    out = {}
    for i in range(E):
        if buf[i] >= 0:
            out[i + lo] = sorted(int(x) for x in buf[2 * E + buf[i]: 2 * E + buf[i] + buf[E + i]])
    return out


def test_perfect_one_to_one_1(oracle_mod):
    # | perfect one-to-one | payloads 0 1 2 3 4 5 6 7 8 9 |     (JoinHashTableTest.cpp:139-150)
    kind, buf, E = _perfect(oracle_mod, [0, 1, 2, 3, 4, 5, 6, 7, 8, 9])
    assert kind == "OneToOne" and buf.tolist() == list(range(10))


def test_perfect_one_to_one_2(oracle_mod):
    # | perfect one-to-one | payloads 0 1 2 * 3 4 5 6 * 7 |     (JoinHashTableTest.cpp:174-184)
    kind, buf, E = _perfect(oracle_mod, [0, 1, 2, 4, 5, 6, 7, 9])
    assert kind == "OneToOne" and buf.tolist() == [0, 1, 2, -1, 3, 4, 5, 6, -1, 7]


def test_perfect_one_to_many_1(oracle_mod):
    # | offsets 0 2 4 6 8 | counts 2 2 2 2 2 | payloads 0 5 1 6 2 7 3 8 4 9 |   (:213-216)
    kind, buf, E = _perfect(oracle_mod, [0, 1, 2, 3, 4, 0, 1, 2, 3, 4])
    assert kind == "OneToMany"
    assert buf[:E].tolist() == [0, 2, 4, 6, 8] and buf[E:2 * E].tolist() == [2, 2, 2, 2, 2]
    assert buf[2 * E:].tolist() == [0, 5, 1, 6, 2, 7, 3, 8, 4, 9]
    assert decode_one_to_many(buf, E) == {0: [0, 5], 1: [1, 6], 2: [2, 7], 3: [3, 8], 4: [4, 9]}


def test_perfect_one_to_many_2(oracle_mod):
    # | offsets 0 * 2 4 6 | counts 2 * 2 2 2 | payloads 0 4 1 5 2 6 3 7 |        (:244-247)
    kind, buf, E = _perfect(oracle_mod, [0, 2, 3, 4, 0, 2, 3, 4])
    assert kind == "OneToMany"
    assert buf[:E].tolist() == [0, -1, 2, 4, 6] and buf[E:2 * E].tolist() == [2, 0, 2, 2, 2]
    assert decode_one_to_many(buf, E) == {0: [0, 4], 2: [1, 5], 3: [2, 6], 4: [3, 7]}


def test_multi_fragment_perfect(oracle_mod):
    # MultiFragment.PerfectOneToOne / PerfectOneToMany (JoinHashTableTest.cpp:444-560): the same
    # tables split into fragments must decode to the same sets, row ids being global
    kind, buf, E = _perfect(oracle_mod, [0, 1, 2, 3, 4, 5, 6, 7, 8, 9], frag=3)
    assert kind == "OneToOne" and buf.tolist() == list(range(10))
    kind, buf, E = _perfect(oracle_mod, [0, 1, 2, 3, 4, 0, 1, 2, 3, 4], frag=4)
    assert decode_one_to_many(buf, E) == {0: [0, 5], 1: [1, 6], 2: [2, 7], 3: [3, 8], 4: [4, 9]}


def _keyed(oracle, b_vals, with_val):
    b = np.array(b_vals, dtype=np.int32)
    E = 2 * len(b)  # BaselineJoinHashTable sizes the table at 2 × the inner row count (JHT/BaselineJoinHashTable.cpp:256-259)
    jc1, jc2 = oracle.make_join_column([b], 4), oracle.make_join_column([b], 4)   # keep the chunk arrays alive
    jcs = (abi.JoinColumn * 2)(jc1, jc2)
    tis = (abi.JoinColumnTypeInfo * 2)(oracle.make_type_info(4, 0, 3, abi.int_null(4)), oracle.make_type_info(4, 0, 3, abi.int_null(4)))
    L = oracle.lib()
    n = 2 + (1 if with_val else 0)
    buf = np.empty(E * n, dtype=np.int32)
    L.oracle_init_baseline_hash_join_buff(buf.ctypes.data, E, 2, int(with_val), -1, 4)
    rc = L.oracle_fill_baseline_hash_join_buff(buf.ctypes.data, E, -1, 0, 2, int(with_val), jcs, tis, 4)
    return rc, buf.reshape(E, n), E


def test_keyed_one_to_one(oracle_mod):
    # | keyed one-to-one | keys * (1,1,1) (3,3,2) (0,0,0) * * |   (JoinHashTableTest.cpp:366-367)
    rc, buf, E = _keyed(oracle_mod, [0, 1, 3], True)
    e = abi.EMPTY_KEY_32
    assert rc == 0
    assert buf.tolist() == [[e, e, -1], [1, 1, 1], [3, 3, 2], [0, 0, 0], [e, e, -1], [e, e, -1]]


def test_keyed_one_to_many_dictionary(oracle_mod):
    # | keyed one-to-many | keys * (1,1) (3,3) (0,0) * * * * |  — the composite-key dictionary of
    # JoinHashTableTest.cpp:411-413 (4 inner rows ⇒ 8 entries; same hash positions modulo 8 are pinned by
    # the reference runtime below).  A duplicate key makes the one-to-one build fail with -1.
    rc, buf, E = _keyed(oracle_mod, [0, 1, 3, 3], True)
    assert rc == -1


@pytest.mark.parametrize("kind", KINDS)
def test_baseline_probe_matches_reference_runtime(oracle_mod, kind):
    L = _lib(oracle_mod, kind)
    rc, buf, E = _keyed(oracle_mod, [0, 1, 3], True)
    keys = np.array([[0, 0], [1, 1], [3, 3], [2, 2], [1, 3]], dtype=np.int32)
    out = np.zeros(len(keys), dtype=np.int64)
    L.oracle_probe_baseline_hash_join(buf.ctypes.data, keys.ctypes.data, len(keys), 2, 4, E, out.ctypes.data)
    assert out[:3].tolist() == [0, 1, 2] and all(x < 0 for x in out[3:])


# ------------------------------------------------------------------ get_group_value semantics ---
@pytest.mark.parametrize("kind", KINDS)
def test_group_by_hash_set_get(oracle_mod, kind):
    """GroupByHashTest.cpp SetGetTest.{OneKey,ManyKeys,OneKeyCollision,OneKeyAllCollisions} + full table ⇒ NULL"""
    L = _lib(oracle_mod, kind)

    def fresh(entries, kq):
        b = np.zeros(entries * (kq + 1), dtype=np.int64)
        b.reshape(entries, kq + 1)[:, :kq] = abi.EMPTY_KEY_64
        return b

    def get(buf, entries, key, kq):
        k = np.array(key, dtype=np.int64)
        p = L.oracle_get_group_value(buf.ctypes.data, entries, k.ctypes.data, kq, 8, kq + 1)
        return None if not p else (p - buf.ctypes.data) // 8

    gb = fresh(10, 1)
    a = get(gb, 10, [31], 1)
    assert a is not None and get(gb, 10, [31], 1) == a
    gb[a] = 42
    assert gb[get(gb, 10, [31], 1)] == 42
    gb = fresh(10, 5)
    a = get(gb, 10, [31, 32, 33, 34, 35], 5)
    assert a is not None and get(gb, 10, [31, 32, 33, 34, 35], 5) == a
    gb = fresh(10, 1)
    a, b = get(gb, 10, [31], 1), get(gb, 10, [41], 1)
    assert a is not None and b is not None and a != b
    gb[a], gb[b] = 32, 42
    assert gb[get(gb, 10, [31], 1)] == 32 and gb[get(gb, 10, [41], 1)] == 42
    # fill the table: 10 distinct keys fit, the 11th gets NULL, existing keys still resolve
    gb = fresh(10, 1)
    slots = [get(gb, 10, [k], 1) for k in range(100, 110)]
    assert None not in slots and len(set(slots)) == 10
    assert get(gb, 10, [999], 1) is None
    assert [get(gb, 10, [k], 1) for k in range(100, 110)] == slots


def test_port_matches_reference_runtime_hashes(oracle_mod):
    if not oracle_mod.ref_available():
        pytest.skip("oracle/_ref not built")
    P, R = oracle_mod.lib("port"), oracle_mod.lib("reference")
    rng = np.random.default_rng(7)
    for n in [1, 3, 4, 7, 8, 12, 16, 31, 64]:
        b = rng.integers(0, 256, n, dtype=np.uint8)
        assert P.oracle_murmur3(b.ctypes.data, n, 0) == R.oracle_murmur3(b.ctypes.data, n, 0)
        assert P.oracle_murmur1(b.ctypes.data, n, 0) == R.oracle_murmur1(b.ctypes.data, n, 0)
        assert P.oracle_murmur64a(b.ctypes.data, n, 0) == R.oracle_murmur64a(b.ctypes.data, n, 0)
    for t in [0, 1, 86399, 86400, 951782400, 951868800, 1230768000, 1467331200, 4102444800, -1, -86400, -2208988800,
              2085978496, 2085978497, 253402300799]:
        assert P.oracle_extract_year(t) == R.oracle_extract_year(t), t
    ts = rng.integers(-10**10, 10**10, 2000)
    assert [P.oracle_extract_year(int(t)) for t in ts] == [R.oracle_extract_year(int(t)) for t in ts]


def test_extract_year_known_answers(oracle_mod):
    L = oracle_mod.lib()
    import datetime
    for y, m, d in [(1970, 1, 1), (1999, 12, 31), (2000, 2, 29), (2009, 1, 1), (2016, 6, 30), (2038, 1, 19), (1900, 3, 1), (2100, 12, 31)]:
        t = int((datetime.datetime(y, m, d, 23, 59, 59) - datetime.datetime(1970, 1, 1)).total_seconds())
        assert L.oracle_extract_year(t) == y


# ------------------------------------------------------------------- end-to-end golden queries ---
@pytest.mark.parametrize("kind", KINDS)
def test_partitioned_group_by_vectors(oracle_mod, kind):
    """PartitionedGroupByTest.cpp:39-139: 20 rows, 5 fragments, keys int64/int32/int16(/dict), sum(v1), sum(v2)."""
    _lib(oracle_mod, kind)
    n = 20
    i = np.arange(1, n + 1)
    id1 = np.where(i == n, 1000000000000, i).astype(np.int64)   # the 1e12 key forces baseline hash
    t = pa.table({"id1": id1, "id2": (i * 10).astype(np.int32), "id3": (i * 100).astype(np.int16),
                  "v1": (i * 3).astype(np.int32), "v2": (i * 111).astype(np.int32)})
    st = util.make_storage({"test1": t}, fragment_size=n // 5)
    pq = util.plan_sql(st, "SELECT id1, SUM(v1) FROM test1 GROUP BY id1", max_groups_buffer_entry_count=64)
    assert pq.qmd.hash_type == abi.BASELINE_HASH
    buf, err = util.run_oracle(oracle_mod, st, pq, kind=kind)
    assert err == 0
    rows = util.sort_rows(util.result_columns(oracle_mod, pq, buf), 1)
    assert rows == sorted(zip(id1.tolist(), (i * 3).tolist()))
    pq = util.plan_sql(st, "SELECT id1, id2, id3, SUM(v1), SUM(v2) FROM test1 GROUP BY id1, id2, id3",
                       max_groups_buffer_entry_count=64)
    buf, err = util.run_oracle(oracle_mod, st, pq, kind=kind)
    rows = util.sort_rows(util.result_columns(oracle_mod, pq, buf), 3)
    assert rows == sorted(zip(id1.tolist(), (i * 10).tolist(), (i * 100).tolist(), (i * 3).tolist(), (i * 111).tolist()))


@pytest.mark.parametrize("kind", KINDS)
def test_pyhdk_api_agg_vectors(oracle_mod, kind):
    """python/tests/test_pyhdk_api.py:457-497"""
    _lib(oracle_mod, kind)
    t = pa.table({"a": [1, 2, 1, 2, 1, 2, 1, 2, 1, 2], "b": [1, 1, 1, 1, 1, 2, 2, 2, 2, 2], "c": list(range(1, 11))})
    st = util.make_storage({"t": t}, fragment_size=4)
    pq = util.plan_sql(st, "SELECT a, b, SUM(c), MIN(c), COUNT(*) FROM t GROUP BY a, b")
    buf, err = util.run_oracle(oracle_mod, st, pq, kind=kind)
    rows = util.sort_rows(util.result_columns(oracle_mod, pq, buf), 2)
    assert rows == [(1, 1, 9, 1, 3), (1, 2, 16, 7, 2), (2, 1, 6, 2, 2), (2, 2, 24, 6, 3)]
    pq = util.plan_sql(st, "SELECT a, COUNT(b), MAX(c), MIN(c), AVG(c) FROM t GROUP BY a")
    buf, err = util.run_oracle(oracle_mod, st, pq, kind=kind)
    rows = util.sort_rows(util.result_columns(oracle_mod, pq, buf), 1)
    assert rows == [(1, 5, 9, 1, 5.0), (2, 5, 10, 2, 6.0)]


@pytest.mark.parametrize("kind", KINDS)
def test_pyhdk_api_join_vectors(oracle_mod, kind):
    """python/tests/test_pyhdk_api.py:609-640: ht1 ⋈ ht2 on a — followed by a group-by over the joined row"""
    _lib(oracle_mod, kind)
    t1 = pa.table({"a": [1, 2, 3, 4, 5], "b": [5, 4, 3, 2, 1], "x": [1.1, 2.2, 3.3, 4.4, 5.5]})
    t2 = pa.table({"a": [1, 2, 3, 4, 5], "b": [1, 2, 3, 4, 5], "y": [5.5, 4.4, 3.3, 2.2, 1.1]})
    st = util.make_storage({"ht1": t1, "ht2": t2}, fragment_size=8)
    pq = util.plan_sql(st, "SELECT ht1.a, SUM(ht1.x), SUM(ht2.y), MAX(ht2.b) FROM ht1 JOIN ht2 ON ht1.a = ht2.a GROUP BY ht1.a")
    buf, err = util.run_oracle(oracle_mod, st, pq, kind=kind)
    rows = util.sort_rows(util.result_columns(oracle_mod, pq, buf), 1)
    util.assert_rows_equal(rows, [(1, 1.1, 5.5, 1), (2, 2.2, 4.4, 2), (3, 3.3, 3.3, 3), (4, 4.4, 2.2, 4), (5, 5.5, 1.1, 5)])
    # join on ht1.a = ht2.b with a partially matching inner side
    t3 = pa.table({"b": [2, 4, 6], "y": [20.0, 40.0, 60.0]})
    st = util.make_storage({"ht1": t1, "ht3": t3}, fragment_size=8)
    pq = util.plan_sql(st, "SELECT ht1.b, COUNT(*), SUM(ht3.y) FROM ht1 JOIN ht3 ON ht1.a = ht3.b GROUP BY ht1.b")
    buf, err = util.run_oracle(oracle_mod, st, pq, kind=kind)
    rows = util.sort_rows(util.result_columns(oracle_mod, pq, buf), 1)
    util.assert_rows_equal(rows, [(2, 1, 40.0), (4, 1, 20.0)])


def test_port_and_reference_buffers_identical(oracle_mod):
    """The restated runtime and the reference's own runtime must leave byte-identical buffers."""
    if not oracle_mod.ref_available():
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(3)
    n = 20000
    t = pa.table({
        "k": rng.integers(-50, 50, n).astype(np.int32), "k2": rng.integers(0, 7, n).astype(np.int16),
        "v": pa.array(rng.integers(-2**40, 2**40, n), mask=rng.random(n) < 0.05),
        "w": pa.array(rng.integers(-1000, 1000, n).astype(np.int32), mask=rng.random(n) < 0.5),
        "f": pa.array(rng.normal(0, 1e3, n), mask=rng.random(n) < 0.1),
        "g": pa.array(rng.normal(0, 10, n).astype(np.float32)),
        "big": rng.integers(-2**62, 2**62, n)})
    st = util.make_storage({"t": t}, fragment_size=3000)
    queries = [
        ("SELECT k, COUNT(*), COUNT(v), SUM(v), MIN(v), MAX(v), AVG(v) FROM t GROUP BY k", {}),
        ("SELECT k, k2, SUM(w), AVG(w), MIN(f), MAX(f), SUM(f), AVG(f) FROM t GROUP BY k, k2", {}),
        ("SELECT k2, SUM(g), MIN(g), MAX(g), AVG(g), COUNT(g) FROM t GROUP BY k2", {}),
        ("SELECT k, SUM(f * (1 - g)), COUNT(*) FROM t WHERE w > 0 AND f < 500 GROUP BY k", {}),
        ("SELECT big, COUNT(*), SUM(w) FROM t GROUP BY big", dict(max_groups_buffer_entry_count=50000)),
        ("SELECT big, k2, MIN(v), MAX(f) FROM t GROUP BY big, k2", dict(max_groups_buffer_entry_count=50000)),
        ("SELECT k, COUNT(*), SUM(v) FROM t GROUP BY k", dict(output_columnar=True)),
        ("SELECT big, SUM(f), COUNT(w) FROM t GROUP BY big", dict(max_groups_buffer_entry_count=50000, output_columnar=True)),
    ]
    for text, kw in queries:
        pq = util.plan_sql(st, text, **kw)
        b1, e1 = util.run_oracle(oracle_mod, st, pq, kind="port")
        b2, e2 = util.run_oracle(oracle_mod, st, pq, kind="reference")
        assert e1 == e2 == 0, text
        assert np.array_equal(b1, b2), text
