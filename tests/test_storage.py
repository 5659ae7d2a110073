"""ArrowStorage façade (hdk_b200/storage.py) against the physical format the reference's ArrowStorage produces
(omniscidb/ArrowStorage/ArrowStorageUtils.cpp:100-170 null conversion, ArrowStorage.cpp:860-1040 fragmenting and chunk
statistics; omniscidb/Tests/ArrowStorageTest.cpp checks the same things): fragment boundaries, NULL → in-band sentinel per
type, date32 as days, timestamp units, dictionary ids shared across fragments, min / max / has_nulls, sharding."""
import datetime

import numpy as np
import pyarrow as pa
import pytest

from hdk_b200 import abi, storage


def make_table(n=1000):
    rng = np.random.default_rng(17)
    return pa.table({
        "i8": pa.array(rng.integers(-100, 100, n).astype(np.int8), mask=rng.random(n) < 0.1),
        "i16": pa.array(rng.integers(-1000, 1000, n).astype(np.int16), mask=rng.random(n) < 0.1),
        "i32": pa.array(rng.integers(-10**6, 10**6, n).astype(np.int32), mask=rng.random(n) < 0.1),
        "i64": pa.array(rng.integers(-2**50, 2**50, n), mask=rng.random(n) < 0.1),
        "f32": pa.array(rng.normal(0, 1, n).astype(np.float32), mask=rng.random(n) < 0.1),
        "f64": pa.array(rng.normal(0, 1, n), mask=rng.random(n) < 0.1),
        "ts": pa.array((rng.integers(0, 10**9, n) * 1000).astype("datetime64[ms]"), mask=rng.random(n) < 0.1),
        "d32": pa.array(rng.integers(0, 20000, n).astype(np.int32), type=pa.int32(), mask=rng.random(n) < 0.1).cast(pa.date32()),
        "s": pa.array(rng.choice(["a", "bb", "ccc", None], n).tolist()),
        "nn": pa.array(rng.integers(0, 5, n).astype(np.int32)),
    })


def test_fragment_boundaries_and_row_offsets():
    t = make_table(1000)
    tab = storage.ArrowStorage().import_arrow_table(t, "t", fragment_size=300)
    assert [f.num_rows for f in tab.fragments] == [300, 300, 300, 100]
    assert [f.row_offset for f in tab.fragments] == [0, 300, 600, 900]
    assert tab.num_rows == 1000
    for f in tab.fragments:
        for c, ci in tab.columns.items():
            assert f.chunks[c].dtype.itemsize == ci.phys_width and len(f.chunks[c]) == f.num_rows


def test_nulls_become_in_band_sentinels():
    t = make_table(500)
    tab = storage.ArrowStorage().import_arrow_table(t, "t", fragment_size=500)
    f = tab.fragments[0]
    for c, sentinel in [("i8", -2**7), ("i16", -2**15), ("i32", -2**31), ("i64", -2**63), ("ts", -2**63), ("d32", -2**31)]:
        mask = np.asarray(t.column(c).is_null())
        assert (f.chunks[c][mask] == sentinel).all() and not (f.chunks[c][~mask] == sentinel).any(), c
    m32, m64 = np.asarray(t.column("f32").is_null()), np.asarray(t.column("f64").is_null())
    assert (f.chunks["f32"][m32] == np.float32(abi.FLT_MIN)).all() and (f.chunks["f64"][m64] == abi.DBL_MIN).all()
    # values survive: date32 stays in days (decoded to seconds inside the kernels), timestamps keep their unit
    ok = ~np.asarray(t.column("d32").is_null())
    assert np.array_equal(f.chunks["d32"][ok], np.asarray(t.column("d32").cast(pa.int32()).fill_null(0))[ok])
    assert tab.columns["d32"].type.date_in_days and tab.columns["d32"].phys_width == 4 and tab.columns["d32"].type.width == 8
    assert tab.columns["ts"].type.unit == 1000


def test_dictionary_ids_are_table_wide():
    t = make_table(900)
    tab = storage.ArrowStorage().import_arrow_table(t, "t", fragment_size=250)
    d = tab.columns["s"].dictionary
    assert sorted(d) == ["a", "bb", "ccc"]
    vals = t.column("s").to_pylist()
    off = 0
    for f in tab.fragments:
        ids = f.chunks["s"]
        for i, v in enumerate(vals[off: off + f.num_rows]):
            assert (ids[i] == abi.int_null(4)) if v is None else (d[ids[i]] == v)
        off += f.num_rows


def test_chunk_statistics_drive_the_planner():
    t = make_table(800)
    tab = storage.ArrowStorage().import_arrow_table(t, "t", fragment_size=256)
    for c in ("i8", "i16", "i32", "i64", "f64"):
        col = t.column(c)
        lo, hi, hn = tab.col_stats(c)
        exp = pa.compute.min_max(col).as_py()
        assert lo == exp["min"] and hi == exp["max"] and hn is True, c
    lo, hi, hn = tab.col_stats("nn")
    assert (lo, hi, hn) == (0, 4, False)
    # a fragment whose values are all NULL has no range but reports has_nulls
    allnull = pa.table({"x": pa.array([None, None, 3, 4], type=pa.int32())})
    tb = storage.ArrowStorage().import_arrow_table(allnull, "n", fragment_size=2)
    s0, s1 = tb.fragments[0].stats["x"], tb.fragments[1].stats["x"]
    assert (s0.min, s0.max, s0.has_nulls) == (None, None, True) and (s1.min, s1.max, s1.has_nulls) == (3, 4, False)
    assert tb.col_stats("x") == (3, 4, True)


def test_sharding_keeps_every_fragment_exactly_once():
    t = make_table(1000)
    kept = []
    for r in range(3):
        tab = storage.ArrowStorage().import_arrow_table(t, "t", fragment_size=128, shard=(r, 3))
        kept += [f.frag_id for f in tab.fragments]
        assert all(f.frag_id % 3 == r for f in tab.fragments)
    assert sorted(kept) == list(range(8))


def test_duplicate_table_and_unsupported_types():
    st = storage.ArrowStorage()
    st.import_arrow_table(pa.table({"a": [1, 2]}), "t")
    with pytest.raises(ValueError):
        st.import_arrow_table(pa.table({"a": [1]}), "t")
    with pytest.raises(NotImplementedError):
        st.import_arrow_table(pa.table({"l": pa.array([[1], [2]])}), "lists")
    st.drop_table("t")
    st.import_arrow_table(pa.table({"a": [datetime.date(2020, 1, 2)]}), "t")
    assert st.get_table("t").fragments[0].chunks["a"][0] == (datetime.date(2020, 1, 2) - datetime.date(1970, 1, 1)).days
