"""Layout decisions of the planner against the reference's rules (QE/MemoryLayoutBuilder.cpp) —
perfect vs baseline hash, keyless, slot widths, init values; mirrors omniscidb/Tests/GroupByTest.cpp:61-228."""
import numpy as np
import pyarrow as pa
import pytest

from hdk_b200 import abi, planner, sql
from tests import util


@pytest.fixture(scope="module")
def st():
    rng = np.random.default_rng(0)
    n = 5000
    t = pa.table({
        "k": rng.integers(0, 1000, n).astype(np.int32),
        "small": rng.integers(0, 10, n).astype(np.int16),
        "wide": rng.integers(0, 2**40, n),
        "v": pa.array(rng.integers(1, 100, n), mask=rng.random(n) < 0.1),
        "pos": rng.integers(1, 100, n),
        "f": rng.uniform(1.0, 2.0, n),
        "ts": pa.array((rng.integers(1230768000, 1467331200, n) * 1000).astype("datetime64[ms]")),
    })
    # Arrow marks every column nullable unless told otherwise; `pos` is declared NOT NULL
    fields = [pa.field(f.name, f.type, nullable=(f.name != "pos")) for f in t.schema]
    t = pa.table([t.column(f.name) for f in fields], schema=pa.schema(fields))
    return util.make_storage({"t": t}, fragment_size=2000)


def test_single_int_key_is_perfect_hash(st):
    pq = util.plan_sql(st, "SELECT k, COUNT(*), SUM(v) FROM t GROUP BY k")
    q = pq.qmd
    lo, hi, _ = st.get_table("t").col_stats("k")
    assert q.hash_type == abi.PERFECT_HASH and q.entry_count == hi - lo + 1 and q.min_val == lo and q.max_val == hi
    assert q.key_width == 8 and q.key_count == 1


def test_keyless_detection_and_target_index(st):
    # COUNT(*) makes the layout keyless; index counts slots before it (MemoryLayoutBuilder.cpp:249-416)
    q = util.plan_sql(st, "SELECT k, COUNT(*), SUM(v) FROM t GROUP BY k").qmd
    assert q.keyless == 1 and q.target_idx_for_key == 1
    q = util.plan_sql(st, "SELECT k, AVG(f), COUNT(*) FROM t GROUP BY k").qmd
    assert q.keyless == 1 and q.target_idx_for_key == 2          # AVG's count slot
    q = util.plan_sql(st, "SELECT k, SUM(pos) FROM t GROUP BY k").qmd
    assert q.keyless == 1 and q.target_idx_for_key == 1          # strictly positive non-null SUM
    q = util.plan_sql(st, "SELECT k, SUM(v) FROM t GROUP BY k").qmd
    assert q.keyless == 0                                        # nullable argument with NULLs
    q = util.plan_sql(st, "SELECT k, COUNT(v) FROM t GROUP BY k").qmd
    assert q.keyless == 0


def test_slot_widths(st):
    # only COUNT(*) and ≤ 4-byte keys ⇒ 4-byte slots (pick_target_compact_width, :559-652)
    q = util.plan_sql(st, "SELECT small, COUNT(*) FROM t GROUP BY small").qmd
    assert list(q.slot_padded)[:2] == [4, 4]
    q = util.plan_sql(st, "SELECT small, COUNT(*), SUM(pos) FROM t GROUP BY small").qmd
    assert list(q.slot_padded)[:3] == [8, 8, 8]
    q = util.plan_sql(st, "SELECT small, k, COUNT(*) FROM t GROUP BY small, k").qmd
    assert list(q.slot_padded)[:3] == [8, 8, 8]                  # multi-key ⇒ crt_min_byte_width
    q = util.plan_sql(st, "SELECT small, AVG(f) FROM t GROUP BY small").qmd
    assert q.slot_count == 3                                     # key, sum, count


def test_init_values(st):
    q = util.plan_sql(st, "SELECT k, COUNT(*), SUM(v), MIN(v), MAX(v), MIN(pos), MAX(pos), SUM(f), MIN(f) FROM t GROUP BY k").qmd
    iv = list(q.init_vals)[:q.slot_count]
    nul = -(1 << 63)
    assert iv[0] == 0 and iv[1] == 0
    assert iv[2] == nul and iv[3] == nul and iv[4] == nul        # nullable int64: the NULL sentinel
    assert iv[5] == (1 << 63) - 1 and iv[6] == nul               # non-null MIN → INT64_MAX, MAX → INT64_MIN
    # the `f` column is declared nullable by Arrow (no NULLs present): sentinel = bits of DBL_MIN
    import struct
    assert iv[7] == struct.unpack("<q", struct.pack("<d", abi.DBL_MIN))[0]


def test_multi_key_perfect_vs_baseline(st):
    q = util.plan_sql(st, "SELECT small, EXTRACT(YEAR FROM ts) AS y, COUNT(*) FROM t GROUP BY small, y").qmd
    assert q.hash_type == abi.PERFECT_HASH and q.entry_count == 10 * 8
    # Π cardinality > baseline_threshold (1e6) ⇒ baseline (MemoryLayoutBuilder.cpp:121-158)
    q = util.plan_sql(st, "SELECT k, wide, COUNT(*) FROM t GROUP BY k, wide").qmd
    assert q.hash_type == abi.BASELINE_HASH and q.entry_count == 16384 and q.key_width == 8
    # wide single key ⇒ baseline; 32-bit range keys compact to 4 bytes (:654-690)
    q = util.plan_sql(st, "SELECT wide, COUNT(*) FROM t GROUP BY wide").qmd
    assert q.hash_type == abi.BASELINE_HASH and q.keyless == 0
    cfg = planner.Config(baseline_threshold=100)
    q = util.plan_sql(st, "SELECT small, k, COUNT(*) FROM t GROUP BY small, k", cfg=cfg).qmd
    assert q.hash_type == abi.BASELINE_HASH and q.key_width == 4
    assert all(x == 8 for x in list(q.slot_padded)[:q.slot_count])


def test_unsupported_shapes_raise(st):
    with pytest.raises(planner.UnsupportedPlan):
        util.plan_sql(st, "SELECT COUNT(*) FROM t GROUP BY f")
    with pytest.raises(planner.UnsupportedPlan):
        sql.parse("SELECT k FROM nosuch GROUP BY k", st.tables)
    with pytest.raises(planner.UnsupportedPlan):
        util.plan_sql(st, "SELECT k, f FROM t GROUP BY k")
