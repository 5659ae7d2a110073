"""Synthetic inputs of the named configs, generated ON THE DEVICE with torch's Philox
generators (one generator per (seed, rank, fragment, column) so every shard is reproducible without
moving data).  Shapes: taxi (omniscidb/Benchmarks/taxi/taxi_full_bench.cpp:41-100 schema), config-1
(10M-row k/v table), TPC-H lineitem columns used by Q1, config-4 composite keys, config-5 star join.
Benchmark / test infrastructure only — not part of the product path."""
from __future__ import annotations

import datetime

import numpy as np
import torch

from hdk_b200 import ir
from hdk_b200.storage import ArrowStorage, ChunkStats, ColumnInfo, Fragment

SEED = 20240917
FRAGMENT_ROWS = 32_000_000      # ArrowStorage default fragment size (ArrowStorage.h:39)


def _gen(device, *key):
    g = torch.Generator(device=device)
    g.manual_seed(hash((SEED,) + tuple(key)) & 0x7FFFFFFFFFFFFFFF)
    return g


def _epoch_ms(y, m, d):
    return int((datetime.datetime(y, m, d) - datetime.datetime(1970, 1, 1)).total_seconds()) * 1000


TAXI_SCHEMA = {
    "cab_type": (ir.SqlType("dict", 4, True, dict_id=1), np.int32),
    "passenger_count": (ir.SqlType("int", 2, True), np.int16),
    "pickup_datetime": (ir.SqlType("timestamp", 8, True, unit=1000), np.int64),
    "total_amount": (ir.SqlType("fp", 8, True), np.float64),
    "trip_distance": (ir.SqlType("fp", 8, True), np.float64),
}
PASSENGER_PMF = [0.0005, 0.70, 0.14, 0.04, 0.02, 0.055, 0.035, 0.004, 0.003, 0.0025]

TAXI_QUERIES = {
    "q1": "SELECT cab_type, count(*) FROM trips GROUP BY cab_type",
    "q2": "SELECT passenger_count, avg(total_amount) FROM trips GROUP BY passenger_count",
    "q3": "SELECT passenger_count, extract(year from pickup_datetime) AS pickup_year, count(*) FROM trips "
          "GROUP BY passenger_count, pickup_year",
    "q4": "SELECT passenger_count, extract(year from pickup_datetime) AS pickup_year, cast(trip_distance as int) AS distance, "
          "count(*) AS the_count FROM trips GROUP BY passenger_count, pickup_year, distance ORDER BY pickup_year, the_count desc",
}
# algorithmic bytes per row: only the columns a query must read once
TAXI_BYTES_PER_ROW = {"q1": 4, "q2": 10, "q3": 10, "q4": 18}


def taxi_fragment(device, rows, rank, frag_id, columns=None):
    cols = {}
    want = columns or list(TAXI_SCHEMA)
    if "cab_type" in want:
        cols["cab_type"] = (torch.rand(rows, device=device, generator=_gen(device, rank, frag_id, 0)) < 0.08).to(torch.int32)
    if "passenger_count" in want:
        cdf = torch.tensor(np.cumsum(PASSENGER_PMF) / np.sum(PASSENGER_PMF), device=device, dtype=torch.float32)
        u = torch.rand(rows, device=device, generator=_gen(device, rank, frag_id, 1))
        cols["passenger_count"] = torch.searchsorted(cdf, u).clamp_(max=9).to(torch.int16)
    if "pickup_datetime" in want:
        lo, hi = _epoch_ms(2009, 1, 1), _epoch_ms(2016, 7, 1)
        cols["pickup_datetime"] = torch.randint(lo, hi, (rows,), device=device, dtype=torch.int64, generator=_gen(device, rank, frag_id, 2))
    if "total_amount" in want:
        x = torch.randn(rows, device=device, dtype=torch.float64, generator=_gen(device, rank, frag_id, 3))
        cols["total_amount"] = x.mul_(10.0).add_(14.0).abs_()
    if "trip_distance" in want:
        x = torch.empty(rows, device=device, dtype=torch.float64)
        x.exponential_(1.0 / 2.9, generator=_gen(device, rank, frag_id, 4))
        cols["trip_distance"] = x.clamp_(max=200.0)
    return cols


def _stats_of(t: torch.Tensor, is_fp: bool):
    if t.numel() == 0:
        return ChunkStats(None, None, False)
    lo, hi = t.min().item(), t.max().item()
    return ChunkStats(float(lo) if is_fp else int(lo), float(hi) if is_fp else int(hi), False)


def register_device_table(storage: ArrowStorage, name, schema, frags_cols, keep_host=False):
    """frags_cols: list of dict col → device tensor.  Chunks stay on the device ("hot")."""
    columns = {c: ColumnInfo(c, t, np.dtype(dt).itemsize, np.dtype(dt), ["yellow", "green"] if t.kind == "dict" else None)
               for c, (t, dt) in schema.items() if c in frags_cols[0]}
    frags, off = [], 0
    for i, cols in enumerate(frags_cols):
        rows = next(iter(cols.values())).numel()
        stats = {c: _stats_of(v, schema[c][0].is_fp) for c, v in cols.items()}
        dev = {c: v.view(torch.uint8) for c, v in cols.items()}
        host = {c: v.cpu().numpy() for c, v in cols.items()} if keep_host else {}
        frags.append(Fragment(i, rows, off, 0, host, stats, dev))
        off += rows
    return storage.add_device_table(name, columns, frags)


def make_taxi(storage: ArrowStorage, device, total_rows, rank=0, fragment_rows=FRAGMENT_ROWS, keep_host=False, name="trips"):
    frags = []
    fid = 0
    for off in range(0, total_rows, fragment_rows):
        rows = min(fragment_rows, total_rows - off)
        frags.append(taxi_fragment(device, rows, rank, fid))
        fid += 1
    return register_device_table(storage, name, TAXI_SCHEMA, frags, keep_host)


# ---- config 1: 10M rows, int32 key with 1K distinct values, int64 / fp64 values ---------------------------
C1_SCHEMA = {"k": (ir.SqlType("int", 4, True), np.int32), "v": (ir.SqlType("int", 8, True), np.int64),
             "f": (ir.SqlType("fp", 8, True), np.float64)}
C1_QUERY = "SELECT k, COUNT(*), SUM(v), MIN(v), MAX(v) FROM c1 GROUP BY k"
C1_QUERY_F = "SELECT k, COUNT(*), SUM(f), MIN(f), MAX(f) FROM c1 GROUP BY k"


def make_c1(storage, device, rows=10_000_000, fragment_rows=2_500_000, keep_host=False, rank=0):
    frags = []
    for fid, off in enumerate(range(0, rows, fragment_rows)):
        n = min(fragment_rows, rows - off)
        frags.append({
            "k": torch.randint(0, 1000, (n,), device=device, dtype=torch.int32, generator=_gen(device, rank, fid, 10)),
            "v": torch.randint(-2**40, 2**40, (n,), device=device, dtype=torch.int64, generator=_gen(device, rank, fid, 11)),
            "f": torch.rand(n, device=device, dtype=torch.float64, generator=_gen(device, rank, fid, 12)).mul_(2e6).sub_(1e6),
        })
    return register_device_table(storage, "c1", C1_SCHEMA, frags, keep_host)


# ---- config 3: TPC-H Q1 columns of lineitem -----------------------------------------------------------------
LINEITEM_SCHEMA = {
    "l_returnflag": (ir.SqlType("dict", 4, True, dict_id=2), np.int32), "l_linestatus": (ir.SqlType("dict", 4, True, dict_id=3), np.int32),
    "l_quantity": (ir.SqlType("fp", 8, True), np.float64), "l_extendedprice": (ir.SqlType("fp", 8, True), np.float64),
    "l_discount": (ir.SqlType("fp", 8, True), np.float64), "l_tax": (ir.SqlType("fp", 8, True), np.float64),
    "l_shipdate": (ir.SqlType("date", 8, True, date_in_days=True), np.int32),
}
TPCH_Q1 = ("SELECT l_returnflag, l_linestatus, sum(l_quantity) AS sum_qty, sum(l_extendedprice) AS sum_base_price, "
           "sum(l_extendedprice * (1 - l_discount)) AS sum_disc_price, "
           "sum(l_extendedprice * (1 - l_discount) * (1 + l_tax)) AS sum_charge, avg(l_quantity) AS avg_qty, "
           "avg(l_extendedprice) AS avg_price, avg(l_discount) AS avg_disc, count(*) AS count_order "
           "FROM lineitem WHERE l_shipdate <= DATE '1998-09-02' GROUP BY l_returnflag, l_linestatus")
TPCH_Q1_BYTES_PER_ROW = 44


TPCH_Q6 = ("SELECT sum(l_extendedprice * l_discount) FROM lineitem WHERE l_shipdate >= DATE '1994-01-01' AND l_shipdate < DATE '1995-01-01' "
           "AND l_discount >= 0.05 AND l_discount <= 0.07 AND l_quantity < 24")
TPCH_Q6_BYTES_PER_ROW = 8 + 8 + 8 + 4   # l_extendedprice, l_discount, l_quantity fp64; l_shipdate date32


def make_lineitem(storage, device, rows, fragment_rows=FRAGMENT_ROWS, keep_host=False, rank=0):
    d0 = (datetime.date(1992, 1, 2) - datetime.date(1970, 1, 1)).days
    d1 = (datetime.date(1998, 12, 1) - datetime.date(1970, 1, 1)).days
    cutoff = (datetime.date(1995, 6, 17) - datetime.date(1970, 1, 1)).days
    frags = []
    for fid, off in enumerate(range(0, rows, fragment_rows)):
        n = min(fragment_rows, rows - off)
        ship = torch.randint(d0, d1 + 1, (n,), device=device, dtype=torch.int32, generator=_gen(device, rank, fid, 20))
        r = torch.rand(n, device=device, generator=_gen(device, rank, fid, 21))
        # dbgen: shipped before the cutoff → linestatus F and returnflag R/A; after → O and N
        old = ship <= cutoff
        flag = torch.where(old, (r < 0.5).to(torch.int32) * 2, torch.ones_like(ship))   # 0 = A, 1 = N, 2 = R
        status = torch.where(old, torch.zeros_like(ship), torch.ones_like(ship))        # 0 = F, 1 = O
        frags.append({
            "l_returnflag": flag.to(torch.int32), "l_linestatus": status.to(torch.int32),
            "l_quantity": torch.randint(1, 51, (n,), device=device, generator=_gen(device, rank, fid, 22)).to(torch.float64),
            "l_extendedprice": torch.rand(n, device=device, dtype=torch.float64, generator=_gen(device, rank, fid, 23)).mul_(104100.0).add_(900.0),
            "l_discount": torch.randint(0, 11, (n,), device=device, generator=_gen(device, rank, fid, 24)).to(torch.float64).div_(100.0),
            "l_tax": torch.randint(0, 9, (n,), device=device, generator=_gen(device, rank, fid, 25)).to(torch.float64).div_(100.0),
            "l_shipdate": ship,
        })
    return register_device_table(storage, "lineitem", LINEITEM_SCHEMA, frags, keep_host)


# ---- config 4: composite (int64, int32) key, 1B rows, ≤ 100M distinct --------------------------------------------
C4_SCHEMA = {"k1": (ir.SqlType("int", 8, True), np.int64), "k2": (ir.SqlType("int", 4, True), np.int32),
             "v": (ir.SqlType("int", 8, True), np.int64)}
C4_QUERY = "SELECT k1, k2, SUM(v), COUNT(*) FROM c4 GROUP BY k1, k2"
C4_BYTES_PER_ROW = 20


def make_c4(storage, device, rows, distinct, fragment_rows=FRAGMENT_ROWS, keep_host=False, rank=0):
    frags = []
    for fid, off in enumerate(range(0, rows, fragment_rows)):
        n = min(fragment_rows, rows - off)
        base = torch.randint(0, distinct, (n,), device=device, dtype=torch.int64, generator=_gen(device, rank, fid, 30))
        k1 = base * 2654435761 + (base % 7919) * (1 << 33)   # spread over the int64 range (a function of `base` only)
        frags.append({"k1": k1, "k2": (base % 1000).to(torch.int32),
                      "v": torch.randint(0, 1000, (n,), device=device, dtype=torch.int64, generator=_gen(device, rank, fid, 31))})
    return register_device_table(storage, "c4", C4_SCHEMA, frags, keep_host)


# ---- config 5: star join fact ⋈ dim ----------------------------------------------------------------------------
FACT_SCHEMA = {"fk": (ir.SqlType("int", 4, True), np.int32), "measure": (ir.SqlType("fp", 8, True), np.float64)}
DIM_SCHEMA = {"pk": (ir.SqlType("int", 4, True), np.int32), "attr": (ir.SqlType("int", 4, True), np.int32)}
C5_QUERY = "SELECT dim.attr, SUM(fact.measure) FROM fact JOIN dim ON fact.fk = dim.pk GROUP BY dim.attr"
C5_BYTES_PER_ROW = 12


def make_star(storage, device, fact_rows, dim_rows, fragment_rows=FRAGMENT_ROWS, keep_host=False, rank=0):
    pk = torch.randperm(dim_rows, device=device, generator=_gen(device, 0, 0, 40)).to(torch.int32)
    register_device_table(storage, "dim", DIM_SCHEMA, [{"pk": pk, "attr": (pk % 1000).to(torch.int32)}], keep_host)
    frags = []
    hi = int(dim_rows * 1.01)   # ~1 % of the foreign keys miss
    for fid, off in enumerate(range(0, fact_rows, fragment_rows)):
        n = min(fragment_rows, fact_rows - off)
        frags.append({"fk": torch.randint(0, hi, (n,), device=device, dtype=torch.int32, generator=_gen(device, rank, fid, 41)),
                      "measure": torch.rand(n, device=device, dtype=torch.float64, generator=_gen(device, rank, fid, 42)).mul_(100.0)})
    return register_device_table(storage, "fact", FACT_SCHEMA, frags, keep_host)
