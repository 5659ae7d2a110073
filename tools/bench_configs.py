#!/usr/bin/env python
"""Device-time measurements of the other named configs (BASELINE.json configs[0], [2], [3], [4]) on one GPU:
scan-kernel time by CUDA events, algorithmic GB/s and fraction of the measured HBM peak.  Not the bench
contract (bench.py is); used to guide optimisation and quoted in DESIGN.md / profiles/.

    python tools/bench_configs.py [--scale 1.0] [--only c1,tpch,c4,c5]
"""
import argparse
import ctypes as C
import json
import os
import sys
import time

import numpy as np  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import benchdata  # noqa: E402
from hdk_b200 import _lib, abi, sql  # noqa: E402
from hdk_b200.executor import Executor  # noqa: E402
from hdk_b200.storage import ArrowStorage  # noqa: E402


def peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return float(json.load(open(p))["hbm_gbs"]) if os.path.exists(p) else 6650.0


ITERS = None   # --iters: total launches per query (profiling runs: 2 = one warm-up + one measured)


def time_query(ex, text, bytes_per_row, rows, reps=5, guess=None, label="", force=None):
    if ITERS:
        reps = max(ITERS - 2, 0)
    unit = sql.parse(text, ex.storage.tables)
    pq = ex.plan(unit, guess)
    prep = ex.prepare(pq)
    L = ex.lib
    st = ex.ctx.stream_ptr()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    times, tot = [], []
    info = None
    ko = None
    if force is not None:   # hdk_b200_debug_set: force an accumulation strategy
        _lib.debug_set("force_strategy", force)
        label += f" [forced strategy {force}]"
    warm = 2 if reps else 1
    for i in range(reps + 2):
        torch.cuda.synchronize()
        e0.record()
        if pq.qmd.hash_type == abi.BASELINE_HASH:
            _lib.check(L.hdk_b200_init_group_by_buffer(C.byref(pq.qmd), prep["out"].data_ptr(), st), "init")
        prep["err"].zero_()
        e1.record()
        info = ex.launch(pq, prep, ko)
        e2.record()
        torch.cuda.synchronize()
        if i >= warm:
            times.append(e1.elapsed_time(e2))
            tot.append(e0.elapsed_time(e2))
    if force is not None:
        _lib.debug_set("force_strategy", -1)
    err = int(prep["err"].item())
    ms = sum(times) / len(times)
    gbs = bytes_per_row * rows / (ms * 1e-3) / 1e9
    res = {"config": label, "rows": rows, "ms": round(ms, 3), "ms_with_init": round(sum(tot) / len(tot), 3), "rows_per_s": rows / (ms * 1e-3),
           "gbs": round(gbs, 1), "frac_of_measured_peak": round(gbs / peak(), 3), "hash": int(pq.qmd.hash_type), "entries": int(pq.qmd.entry_count),
           "strategy": int(info.strategy), "variant": int(info.variant), "grid": int(info.grid), "smem": int(info.smem_bytes), "block": int(info.block), "tile_rows": int(info.tile_rows), "err": err,
           "buffer_mb": round(prep["out"].numel() / 1e6, 1)}
    print(json.dumps(res), flush=True)
    return res


def time_result_side(ex, text, guess=None, label="", limit=10):
    """The result side of a large group-by on the device: compaction of the non-empty entries
    (hdk_b200_compact_result), ORDER BY … LIMIT (hdk_b200_sort_permutation + hdk_b200_gather_rows), D2H of the answer —
    against copying the whole group-by buffer to the host, which is where the reference's ResultSet iteration and sort start."""
    unit = sql.parse(text, ex.storage.tables)
    pq = ex.plan(unit, guess)
    prep = ex.prepare(pq)
    _lib.check(ex.lib.hdk_b200_init_group_by_buffer(C.byref(pq.qmd), prep["out"].data_ptr(), ex.ctx.stream_ptr()), "init")
    prep["err"].zero_()
    ex.launch(pq, prep)
    torch.cuda.synchronize()
    assert int(prep["err"].item()) == 0
    from hdk_b200.executor import ResultSet
    order = ResultSet(pq, np.zeros(0, dtype=np.uint8), {}).order_entries()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    best = None
    for _ in range(3):
        torch.cuda.synchronize()
        ev[0].record()
        cols, n = ex.compact_on_device(pq, prep["out"], to_host=False)
        ev[1].record()
        top = ex.sort_on_device(cols, n, order, limit)
        ev[2].record()
        host = top.cpu()
        ev[3].record()
        torch.cuda.synchronize()
        t = [ev[i].elapsed_time(ev[i + 1]) for i in range(3)]
        best = t if best is None or sum(t) < sum(best) else best
        del cols, top
    pinned = torch.empty(prep["out"].numel(), dtype=torch.uint8, pin_memory=True)
    torch.cuda.synchronize()
    ev[0].record()
    pinned.copy_(prep["out"], non_blocking=True)
    ev[1].record()
    torch.cuda.synchronize()
    res = {"config": label, "groups": n, "limit": limit, "compact_ms": round(best[0], 3), "sort_limit_ms": round(best[1], 3),
           "d2h_answer_ms": round(best[2], 3), "result_side_total_ms": round(sum(best), 3),
           "d2h_whole_buffer_ms": round(ev[0].elapsed_time(ev[1]), 3), "buffer_mb": round(prep["out"].numel() / 1e6, 1),
           "first_row": [int(x) for x in host[:, 0].tolist()]}
    print(json.dumps(res), flush=True)
    return res


def cpu_reference(make, text, rows, guess=None, label=""):
    """The reference's CPU execution shape on a host-resident slice of the same generator (oracle/_ref when present):
    one kernel per fragment with a private buffer on all host cores + reduce.  Reported beside the GPU number.
    (Same role as bench.py's cpu_baseline leg: the oracle is the thing timed as the CPU baseline, never part of the
    GPU path being measured.)"""
    from oracle import oracle
    from tests import util
    st = ArrowStorage()
    make(st, rows)
    pq = util.plan_sql(st, text, **({"max_groups_buffer_entry_count": guess} if guess else {}))
    kind = "reference" if oracle.ref_available() else "port"
    threads = os.cpu_count() or 1
    frs, jt, ic = util.oracle_inputs(oracle, st, pq)
    oracle.run_query(pq, frs, jt, ic, n_threads=threads, kind=kind)        # warm-up (page faults)
    t0 = time.perf_counter()
    _, err = oracle.run_query(pq, frs, jt, ic, n_threads=threads, kind=kind)
    dt = time.perf_counter() - t0
    res = {"config": label + " [CPU " + kind + "]", "rows": rows, "ms": round(dt * 1e3, 1), "rows_per_s": rows / dt, "cores": threads, "err": int(err)}
    print(json.dumps(res), flush=True)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--only", default="c1,tpch,c5,c4")
    ap.add_argument("--force", action="store_true", help="also time the GLOBAL strategy where it is an alternative")
    ap.add_argument("--cpu", action="store_true", help="also time the reference's CPU path (oracle/_ref) on a host slice of each config")
    ap.add_argument("--cpu-rows", type=int, default=8_000_000)
    ap.add_argument("--pa-partitions", type=int, default=0, help="force the partition count of the partitioned aggregation (0 = auto)")
    ap.add_argument("--iters", type=int, default=0, help="total launches per query (0 = the per-config defaults); 2 for ncu captures")
    args = ap.parse_args()
    global ITERS
    ITERS = args.iters
    if args.pa_partitions:
        _lib.debug_set("partitioned_partitions", args.pa_partitions)
    dev = torch.device("cuda", 0)
    only = args.only.split(",")
    out = []
    if "c1" in only:
        st = ArrowStorage()
        benchdata.make_c1(st, dev)
        ex = Executor(st)
        out.append(time_query(ex, benchdata.C1_QUERY, 12, 10_000_000, reps=20, label="c1 int64 (10M rows, 1K groups)"))
        out.append(time_query(ex, benchdata.C1_QUERY_F, 12, 10_000_000, reps=20, label="c1 fp64"))
        if args.cpu:
            cpu_reference(lambda s, n: benchdata.make_c1(s, dev, rows=10_000_000, keep_host=True), benchdata.C1_QUERY, 10_000_000, label="c1 int64")
        if args.force:
            out.append(time_query(ex, benchdata.C1_QUERY, 12, 10_000_000, reps=20, label="c1 int64", force=2))
        del ex, st
        torch.cuda.empty_cache()
    if "tpch" in only:
        rows = int(600_037_902 * args.scale)
        st = ArrowStorage()
        benchdata.make_lineitem(st, dev, rows)
        ex = Executor(st)
        out.append(time_query(ex, benchdata.TPCH_Q1, benchdata.TPCH_Q1_BYTES_PER_ROW, rows, label="tpch q1 SF100 lineitem"))
        out.append(time_query(ex, benchdata.TPCH_Q6, benchdata.TPCH_Q6_BYTES_PER_ROW, rows, label="tpch q6 (non-grouped aggregate)"))
        if args.cpu:
            mk = lambda s, n: benchdata.make_lineitem(s, dev, n, fragment_rows=1_000_000, keep_host=True)  # noqa: E731
            cpu_reference(mk, benchdata.TPCH_Q1, args.cpu_rows, label="tpch q1")
            cpu_reference(mk, benchdata.TPCH_Q6, args.cpu_rows, label="tpch q6")
        del ex, st
        torch.cuda.empty_cache()
    if "c5" in only:
        rows = int(2_000_000_000 * args.scale)
        st = ArrowStorage()
        benchdata.make_star(st, dev, rows, 10_000_000)
        ex = Executor(st)
        t0 = time.perf_counter()
        ex.build_join_table(st.get_table("dim"), "pk")
        torch.cuda.synchronize()
        print(json.dumps({"config": "c5 join build (10M rows)", "ms_first_call": round((time.perf_counter() - t0) * 1e3, 2)}), flush=True)
        out.append(time_query(ex, benchdata.C5_QUERY, benchdata.C5_BYTES_PER_ROW, rows, label="c5 star join 2B x 10M + group-by SUM"))
        if args.force:
            out.append(time_query(ex, benchdata.C5_QUERY, benchdata.C5_BYTES_PER_ROW, rows, label="c5", force=2))
        if args.cpu:
            cpu_reference(lambda s, n: benchdata.make_star(s, dev, n, 10_000_000, fragment_rows=1_000_000, keep_host=True), benchdata.C5_QUERY,
                          args.cpu_rows, label="c5 star join")
        del ex, st
        torch.cuda.empty_cache()
    if "c4" in only:
        rows = int(1_000_000_000 * args.scale)
        distinct = int(100_000_000 * args.scale)
        st = ArrowStorage()
        benchdata.make_c4(st, dev, rows, distinct)
        ex = Executor(st)
        out.append(time_query(ex, benchdata.C4_QUERY, benchdata.C4_BYTES_PER_ROW, rows, reps=3, guess=2 * distinct,
                              label="c4 baseline hash 1B rows / 100M groups"))
        out.append(time_result_side(ex, "SELECT k1, k2, SUM(v) AS s, COUNT(*) AS n FROM c4 GROUP BY k1, k2 ORDER BY s DESC, k1 LIMIT 10",
                                    guess=2 * distinct, label="c4 result side: compact + ORDER BY s DESC, k1 LIMIT 10"))
        if args.cpu:
            cpu_reference(lambda s, n: benchdata.make_c4(s, dev, n, n // 10, fragment_rows=1_000_000, keep_host=True), benchdata.C4_QUERY,
                          args.cpu_rows, guess=2 * (args.cpu_rows // 10), label="c4 baseline hash")
    return 0


if __name__ == "__main__":
    sys.exit(main())
