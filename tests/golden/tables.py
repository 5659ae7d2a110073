"""Deterministic input tables of the golden fixtures: every column derives from `Generator.integers` draws
(bit-identical on every platform) and exact arithmetic, so only the expected outputs need to be stored."""
import numpy as np
import pyarrow as pa

SEED = 20240917
ROWS = 6007


def golden_tables():
    rng = np.random.default_rng(SEED)
    n = ROWS
    ints = lambda lo, hi, dt=np.int64: rng.integers(lo, hi, n).astype(dt)        # noqa: E731
    mask = lambda pct: rng.integers(0, 100, n) < pct                             # noqa: E731
    t = pa.table({
        "k": ints(0, 1000, np.int32),
        "k_null": pa.array(ints(-20, 20, np.int32), mask=mask(3)),
        "s": ints(0, 10, np.int16),
        "b": ints(-3, 3, np.int8),
        "v": pa.array(ints(-2**40, 2**40), mask=mask(1)),
        "w": pa.array(ints(-1000, 1000, np.int32), mask=mask(40)),
        "f": ints(-10**9, 10**9).astype(np.float64) / 1024.0,
        "fn": pa.array(ints(-10**6, 10**6).astype(np.float64) / 64.0, mask=mask(10)),
        "g": pa.array((ints(-10**5, 10**5).astype(np.float64) / 128.0).astype(np.float32), mask=mask(5)),
        "ts": pa.array((ints(1230768000, 1467331200) * 1000 + ints(0, 1000)).astype("datetime64[ms]")),
        "dt": pa.array(ints(8000, 11000, np.int32), type=pa.int32()).cast(pa.date32()),
        "d": ints(0, 200 * 256).astype(np.float64) / 256.0,
        "big": ints(-2**62, 2**62),
        "mid": (np.arange(n) * 7919 % 3001).astype(np.int64) * 1000003,
        "fk": ints(-5, 1010, np.int32),
    })
    pk = rng.permutation(1000).astype(np.int32)
    dim = pa.table({"pk": pk, "attr": (pk % 37).astype(np.int32), "weight": rng.integers(0, 1024, 1000).astype(np.float64) / 1024.0})
    # dimension keyed by the pair (s, b): 10 x 6 combinations, 48 of them present
    combos = rng.permutation(60)[:48]
    dim2k = pa.table({"s": (combos // 6).astype(np.int16), "b": (combos % 6 - 3).astype(np.int8),
                      "attr": rng.integers(0, 5, 48).astype(np.int32), "w": rng.integers(0, 1000, 48)})
    return {"t": t, "dim": dim, "dim2k": dim2k}


FRAGMENT_SIZE = {"t": 1501, "dim": 100000, "dim2k": 100000}

# (name, sql, number of key columns, planner kwargs)
QUERIES = [
    ("c1_int64", "SELECT k, COUNT(*), SUM(v), MIN(v), MAX(v) FROM t GROUP BY k", 1, {}),
    ("c1_fp64", "SELECT k, COUNT(*), SUM(f), MIN(f), MAX(f) FROM t GROUP BY k", 1, {}),
    ("taxi_q1", "SELECT s, COUNT(*) FROM t GROUP BY s", 1, {}),
    ("taxi_q2", "SELECT s, AVG(f) FROM t GROUP BY s", 1, {}),
    ("taxi_q3", "SELECT s, EXTRACT(YEAR FROM ts) AS y, COUNT(*) FROM t GROUP BY s, y", 2, {}),
    ("taxi_q4", "SELECT s, EXTRACT(YEAR FROM ts) AS y, CAST(d AS INT) AS dist, COUNT(*) FROM t GROUP BY s, y, dist", 3, {}),
    ("tpch_q1", "SELECT b, s, SUM(d), SUM(f), SUM(f * (1 - d / 200)), SUM(f * (1 - d / 200) * (1 + d / 100)), AVG(d), AVG(f), AVG(fn), COUNT(*) "
                "FROM t WHERE dt <= DATE '1998-09-02' GROUP BY b, s", 2, {}),
    ("nullable_mix", "SELECT k_null, COUNT(*), COUNT(w), SUM(w), MIN(w), MAX(w), AVG(w), MIN(fn), MAX(fn), SUM(fn) FROM t GROUP BY k_null", 1, {}),
    ("fp32_aggs", "SELECT s, SUM(g), MIN(g), MAX(g), AVG(g), COUNT(g) FROM t GROUP BY s", 1, {}),
    ("filter_logic", "SELECT k, SUM(w), COUNT(*) FROM t WHERE f > 0 AND (s < 5 OR w IS NULL) GROUP BY k", 1, {}),
    ("columnar", "SELECT k_null, s, MIN(v), MAX(fn), AVG(w) FROM t GROUP BY k_null, s", 2, dict(output_columnar=True)),
    ("baseline_1key", "SELECT big, COUNT(*), SUM(w), MIN(fn), MAX(f), AVG(v) FROM t GROUP BY big", 1, dict(max_groups_buffer_entry_count=16384)),
    ("baseline_2key", "SELECT mid, s, COUNT(*), SUM(v), SUM(f) FROM t GROUP BY mid, s", 2, dict(max_groups_buffer_entry_count=32768)),
    ("star_join", "SELECT dim.attr, SUM(t.f), COUNT(*) FROM t JOIN dim ON t.fk = dim.pk GROUP BY dim.attr", 1, {}),
    ("composite_join", "SELECT dim2k.attr, COUNT(*), SUM(t.f), SUM(dim2k.w), MIN(t.v) FROM t JOIN dim2k ON t.s = dim2k.s AND t.b = dim2k.b "
                       "GROUP BY dim2k.attr", 1, {}),
    ("join_filter", "SELECT dim.attr, t.s, SUM(t.f * dim.weight), MIN(t.v) FROM t JOIN dim ON t.fk = dim.pk WHERE dim.weight > 0.25 GROUP BY dim.attr, t.s", 2, {}),
]
