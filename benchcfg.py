"""The other named configs of BASELINE.json (configs[0], [2], [3], [4]) for bench.py's `per_config` block: device time of
the hot path by CUDA events, algorithmic GB/s against the measured HBM peak, a size-independent parity check of the result
at full size, and the reference's CPU runtime on a bounded host sample of the same generator.  Benchmark infrastructure —
not part of the product path (the CPU legs are the only place that touches oracle/)."""
from __future__ import annotations

import ctypes as C
import json
import os
import time

import numpy as np
import torch

import benchdata
from hdk_b200 import _lib, abi, sql
from hdk_b200.executor import Executor
from hdk_b200.storage import ArrowStorage

ROOT = os.path.dirname(os.path.abspath(__file__))


def _traffic(name):
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(p)).get(name)
    except Exception:
        return None


def time_launch(ex: Executor, text: str, reps: int, guess=None):
    """(planned query, prep, avg ms of init-buffer + launch, avg ms of the launch alone, launch info); data resident in HBM."""
    unit = sql.parse(text, ex.storage.tables)
    pq = ex.plan(unit, guess)
    prep = ex.prepare(pq)
    L, st = ex.lib, ex.ctx.stream_ptr()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    t_all, t_launch, info = [], [], None
    for i in range(reps + 3):
        torch.cuda.synchronize()
        e0.record()
        if pq.qmd.hash_type == abi.BASELINE_HASH:     # (Executor.launch initialises the buffer itself: timed separately here)
            _lib.check(L.hdk_b200_init_group_by_buffer(C.byref(pq.qmd), prep["out"].data_ptr(), st), "init")
        prep["err"].zero_()
        e1.record()
        info = abi.LaunchInfo()
        _lib.check(L.hdk_b200_launch(C.byref(pq.plan), C.byref(pq.qmd), None, C.byref(prep["kp"]), prep["scratch"].data_ptr(),
                                     prep["scratch_bytes"], st, C.byref(info)), "launch")
        e2.record()
        torch.cuda.synchronize()
        if i >= 3:
            t_all.append(e0.elapsed_time(e2))
            t_launch.append(e1.elapsed_time(e2))
    assert int(prep["err"].item()) == 0, f"in-band error {int(prep['err'].item())}"
    return pq, prep, sum(t_all) / len(t_all), sum(t_launch) / len(t_launch), info


def _entry(name, rows, bytes_per_row, ms_launch, ms_all, peak, info, pq, parity, extra=None):
    gbs = bytes_per_row * rows / (ms_launch * 1e-3) / 1e9
    d = {"rows": rows, "bytes_per_row": bytes_per_row, "algorithmic_bytes": bytes_per_row * rows, "launch_ms": ms_launch,
         "launch_ms_with_buffer_init": ms_all, "rows_per_s": rows / (ms_launch * 1e-3), "achieved_gbs": gbs, "frac": gbs / peak,
         "traffic": _traffic(name), "strategy": int(info.strategy), "precompiled_shape": int(info.variant) > 0,
         "kernels_per_launch": int(info.n_launches), "entries": int(pq.qmd.entry_count), "parity_check": parity}
    if extra:
        d.update(extra)
    return d


def _compact(ex, pq, prep):
    cols, n = ex.compact_on_device(pq, prep["out"], to_host=False)
    return cols[:, :n], n


# ---- parity at full size: properties that do not need the oracle -------------------------------------------
def check_c1(ex, pq, prep, st, col):
    tab = st.get_table("c1")
    k = torch.cat([f.device_chunks["k"].view(torch.int32) for f in tab.fragments]).to(torch.int64)
    cols, n = _compact(ex, pq, prep)
    order = torch.argsort(cols[0])
    keys, cnt = cols[0][order], cols[1][order]
    exp_cnt = torch.bincount(k, minlength=1000)
    assert n == int((exp_cnt > 0).sum()) and torch.equal(keys, torch.nonzero(exp_cnt).flatten()), "group keys"
    assert torch.equal(cnt, exp_cnt[exp_cnt > 0]), "COUNT(*) per group"
    if col == "v":
        v = torch.cat([f.device_chunks["v"].view(torch.int64) for f in tab.fragments])
        exp_sum = torch.zeros(1000, dtype=torch.int64, device=v.device).scatter_add_(0, k, v)
        assert torch.equal(cols[2][order], exp_sum[exp_cnt > 0]), "SUM(v) per group (bit-exact)"
        exp_min = torch.full((1000,), 2**62, dtype=torch.int64, device=v.device).scatter_reduce_(0, k, v, "amin")
        exp_max = torch.full((1000,), -2**62, dtype=torch.int64, device=v.device).scatter_reduce_(0, k, v, "amax")
        assert torch.equal(cols[3][order], exp_min[exp_cnt > 0]) and torch.equal(cols[4][order], exp_max[exp_cnt > 0]), "MIN / MAX per group"
    else:
        f64 = torch.cat([f.device_chunks["f"].view(torch.float64) for f in tab.fragments])
        exp_sum = torch.zeros(1000, dtype=torch.float64, device=f64.device).scatter_add_(0, k, f64)
        got = cols[2][order].view(torch.float64)
        assert torch.allclose(got, exp_sum[exp_cnt > 0], rtol=1e-9, atol=0), "SUM(f) per group within 1e-9"
    return "ok: keys, COUNT, SUM, MIN, MAX of every group against independent torch reductions of the same columns"


def check_tpch_q1(ex, pq, prep, st):
    import datetime
    tab = st.get_table("lineitem")
    cutoff = (datetime.date(1998, 9, 2) - datetime.date(1970, 1, 1)).days
    n_pass, qty, n_all = 0, 0.0, 0
    for f in tab.fragments:
        ship = f.device_chunks["l_shipdate"].view(torch.int32)
        m = ship <= cutoff
        n_pass += int(m.sum())
        qty += float(f.device_chunks["l_quantity"].view(torch.float64)[m].sum())
        n_all += ship.numel()
    cols, n = _compact(ex, pq, prep)
    assert 3 <= n <= 4, f"{n} groups (A/F, N/O, R/F of benchdata.make_lineitem; dbgen also has N/F)"
    cnt = int(cols[9].sum())
    assert cnt == n_pass, f"sum of count_order {cnt} != rows passing the filter {n_pass}"
    got_qty = float(cols[2].view(torch.float64).sum())
    assert abs(got_qty - qty) <= 1e-9 * abs(qty), "sum of sum_qty"
    avg_qty = cols[6].view(torch.float64)
    assert torch.allclose(avg_qty, cols[2].view(torch.float64) / cols[9].to(torch.float64), rtol=1e-12), "AVG = SUM / COUNT"
    return f"ok: {n} groups, sum(count_order) == {n_pass} rows passing the filter of {n_all}, sum(sum_qty) within 1e-9, AVG == SUM / COUNT"


def check_c4(ex, pq, prep, st, distinct):
    """checksum of checksums: sum over groups of key x COUNT == sum over rows of key (mod 2^64), for both key components;
    sum of COUNT == rows; sum of SUM(v) == sum of v (bit-exact integers); no key twice."""
    tab = st.get_table("c4")
    rows = tab.num_rows
    s_k1 = s_k2 = s_v = 0
    for f in tab.fragments:
        s_k1 = (s_k1 + int(f.device_chunks["k1"].view(torch.int64).sum())) & (2**64 - 1)      # torch sums wrap modulo 2^64
        s_k2 += int(f.device_chunks["k2"].view(torch.int32).sum(dtype=torch.int64))
        s_v += int(f.device_chunks["v"].view(torch.int64).sum())
    cols, n = _compact(ex, pq, prep)
    k1, k2, sv, cnt = cols[0], cols[1], cols[2], cols[3]
    assert int(cnt.sum()) == rows, "sum of COUNT(*)"
    assert int(sv.sum()) == s_v, "sum of SUM(v)"
    assert (int((k1 * cnt).sum()) & (2**64 - 1)) == s_k1, "sum of k1 x COUNT (mod 2^64)"
    assert int((k2 * cnt).sum()) == s_k2, "sum of k2 x COUNT"
    assert n <= distinct and n >= min(distinct, rows) * 0.99, f"{n} groups for {distinct} distinct keys"
    # k1 is a function of the key (benchdata.make_c4): the number of distinct k1 values equals the number of groups
    assert int(torch.unique(k1).numel()) == n, "a key appears in two entries"
    return f"ok: {n} groups, no key twice; sum COUNT == {rows} rows; sum SUM(v), sum k1 x COUNT (mod 2^64), sum k2 x COUNT equal the columns' own sums"


def check_c5(ex, pq, prep, st, dim_rows):
    tab = st.get_table("fact")
    exp = torch.zeros(1000, dtype=torch.float64, device=ex.ctx.device)
    matched = 0
    for f in tab.fragments:
        fk = f.device_chunks["fk"].view(torch.int32).to(torch.int64)
        m = fk < dim_rows                                   # dim.pk is a permutation of 0 .. dim_rows - 1, attr = pk % 1000
        exp.scatter_add_(0, fk[m] % 1000, f.device_chunks["measure"].view(torch.float64)[m])
        matched += int(m.sum())
    cols, n = _compact(ex, pq, prep)
    order = torch.argsort(cols[0])
    assert n == 1000 and torch.equal(cols[0][order], torch.arange(1000, device=ex.ctx.device)), "group keys"
    got = cols[1][order].view(torch.float64)
    assert torch.allclose(got, exp, rtol=1e-9, atol=0), "SUM(measure) per dim.attr within 1e-9"
    return f"ok: 1000 groups; SUM(measure) per group within 1e-9 of a torch scatter-add over the {matched} matching rows"


# ---- CPU leg: the reference's runtime on a host sample of the same generator ----------------------------------------------
def cpu_sample(make, text, rows, guess=None):
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
    assert err == 0
    return {"value": rows / dt, "unit": "rows/s", "cores": threads, "kind": kind,
            "sample": f"{rows} rows of the same generator, one kernel per fragment on {threads} threads + reduce ({dt:.2f} s)"}


def per_config_single_gpu(device, peak, reps=5, cpu=True, only=("c1", "c3", "c4", "c5"), scale=1.0, cpu_rows=8_000_000):
    out = {}
    if "c1" in only:
        st = ArrowStorage()
        benchdata.make_c1(st, device)
        ex = Executor(st, device=device.index)
        for name, text, col in (("c1_int64", benchdata.C1_QUERY, "v"), ("c1_fp64", benchdata.C1_QUERY_F, "f")):
            pq, prep, ms_all, ms, info = time_launch(ex, text, 20)
            out[name] = _entry(name, 10_000_000, 12, ms, ms_all, peak, info, pq, check_c1(ex, pq, prep, st, col),
                               {"workload": "BASELINE.json configs[0]: 10 M rows, int32 key with 1 K distinct values: " + text,
                                "note": "not HBM-bound: 120 MB of input is 18 µs at the HBM peak; the scan kernel sits on the SM's shared-memory pipe "
                                        "(COUNT + SUM + MIN + MAX into 1000 bins = 40 shared wavefronts per 32 rows = 43 µs of pipe time per launch, "
                                        "profiles/README.md round 2) + three launches (init, scan, finalize)"})
        if cpu:
            out["c1_int64"]["cpu_baseline"] = cpu_sample(lambda s, n: benchdata.make_c1(s, device, rows=n, keep_host=True), benchdata.C1_QUERY, 10_000_000)
        del ex, st
        torch.cuda.empty_cache()
    if "c3" in only:
        rows = int(600_037_902 * scale)
        st = ArrowStorage()
        benchdata.make_lineitem(st, device, rows)
        ex = Executor(st, device=device.index)
        pq, prep, ms_all, ms, info = time_launch(ex, benchdata.TPCH_Q1, reps)
        out["tpch_q1"] = _entry("tpch_q1", rows, benchdata.TPCH_Q1_BYTES_PER_ROW, ms, ms_all, peak, info, pq, check_tpch_q1(ex, pq, prep, st),
                                {"workload": "BASELINE.json configs[2]: TPC-H Q1 on synthetic SF100 lineitem (600,037,902 rows x scale)"})
        if cpu:
            out["tpch_q1"]["cpu_baseline"] = cpu_sample(lambda s, n: benchdata.make_lineitem(s, device, n, fragment_rows=1_000_000, keep_host=True),
                                                        benchdata.TPCH_Q1, cpu_rows)
        del ex, st, prep
        torch.cuda.empty_cache()
    if "c5" in only:
        rows, dim_rows = int(2_000_000_000 * scale), 10_000_000
        st = ArrowStorage()
        benchdata.make_star(st, device, rows, dim_rows)
        ex = Executor(st, device=device.index)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ex.build_join_table(st.get_table("dim"), "pk")
        torch.cuda.synchronize()
        build_ms = (time.perf_counter() - t0) * 1e3
        pq, prep, ms_all, ms, info = time_launch(ex, benchdata.C5_QUERY, reps)
        out["c5_star_join"] = _entry("c5_star_join", rows, benchdata.C5_BYTES_PER_ROW, ms, ms_all, peak, info, pq, check_c5(ex, pq, prep, st, dim_rows),
                                     {"workload": "BASELINE.json configs[4]: 2 B-row fact x 10 M-row dimension star join + group-by SUM (x scale)",
                                      "join_build_ms_first_call": build_ms, "dim_rows": dim_rows,
                                      "second_roof": {"bound": "l1tex request rate (one random gather per row = 1.0 SM cycle per row, measured: "
                                                               "profiles/r2_gather_accum.log)", "gather_only_floor_ms": rows / (148 * 1.965e9) * 1e3,
                                                      "frac_of_gather_roof": rows / (148 * 1.965e9) * 1e3 / ms},
                                      "note": "not HBM-bound: ncu counts 1.02 gather sectors + 0.92 shared wavefronts per row through the same "
                                              "L1TEX pipe (profiles/README.md round 2)"})
        if cpu:
            out["c5_star_join"]["cpu_baseline"] = cpu_sample(lambda s, n: benchdata.make_star(s, device, n, dim_rows, fragment_rows=1_000_000, keep_host=True),
                                                             benchdata.C5_QUERY, cpu_rows)
        del ex, st, prep
        torch.cuda.empty_cache()
    if "c4" in only:
        rows, distinct = int(1_000_000_000 * scale), int(100_000_000 * scale)
        st = ArrowStorage()
        benchdata.make_c4(st, device, rows, distinct)
        ex = Executor(st, device=device.index)
        pq, prep, ms_all, ms, info = time_launch(ex, benchdata.C4_QUERY, max(2, reps // 2), guess=2 * distinct)
        out["c4_baseline_hash"] = _entry("c4_baseline_hash", rows, benchdata.C4_BYTES_PER_ROW, ms, ms_all, peak, info, pq,
                                         check_c4(ex, pq, prep, st, distinct),
                                         {"workload": "BASELINE.json configs[3]: 1 B rows, composite (int64, int32) key, 100 M distinct (x scale)",
                                          "distinct": distinct, "group_by_buffer_bytes": int(prep["out"].numel()),
                                          "path": "radix-partitioned aggregation (count, two scatter levels, per-partition aggregation in shared memory)"
                                          if int(info.strategy) == abi.STRATEGY_PARTITIONED else "global-table probe"})
        if cpu:
            out["c4_baseline_hash"]["cpu_baseline"] = cpu_sample(lambda s, n: benchdata.make_c4(s, device, n, n // 10, fragment_rows=1_000_000, keep_host=True),
                                                                 benchdata.C4_QUERY, cpu_rows, guess=2 * (cpu_rows // 10))
        del ex, st, prep
        torch.cuda.empty_cache()
    return out


# ---- N > 1: the configs with a real exchange step (one process per GPU) ------------------------------------------------------
def _max_over_ranks(x, device):
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def _sum_over_ranks(x, device, dtype=torch.int64):
    import torch.distributed as dist
    t = torch.tensor([x], dtype=dtype, device=device)
    dist.all_reduce(t)
    return t.item()


def multi_c5(device, rank, world, rows_per_gpu, reps=3):
    """Star join, fact rows sharded, dimension table built once on rank 0 and broadcast (vs rebuilt on every rank, what the
    reference does per device), partial tables merged inside the kernels over peer memory."""
    import torch.distributed as dist
    from hdk_b200 import distributed as D
    dim_rows = 10_000_000
    st = ArrowStorage()
    benchdata.make_star(st, device, rows_per_gpu, dim_rows, rank=rank)
    st.get_table("fact").shard = (rank, world)
    ex = Executor(st, device=device.index)
    unit = sql.parse(benchdata.C5_QUERY, st.tables)
    build = {}
    for mode in ("rebuild_on_every_rank", "build_once_broadcast"):
        ex.broadcast_join_build = mode == "build_once_broadcast"
        ts = []
        for i in range(3):
            ex.join_tables.clear()
            dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            pq = ex.plan(unit)
            prep = ex.prepare(pq)          # join table + slot-ordered payload
            torch.cuda.synchronize()
            ts.append((time.perf_counter() - t0) * 1e3)
        build[mode + "_ms"] = _max_over_ranks(min(ts), device)
    xchg = ex._exchange_for(pq)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for i in range(reps + 2):
        dist.barrier()
        torch.cuda.synchronize()
        e0.record()
        if xchg is not None:
            ex.launch_exchange(pq, prep, xchg)
        else:
            ex.execute_sharded(pq, prep)
        e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            ts.append(e0.elapsed_time(e1))
    assert ex._agree_on_error(prep["err"]) == 0
    ms = _max_over_ranks(sum(ts) / len(ts), device)
    # parity: every rank holds the same merged table; its sums equal the sum over all ranks of independent torch reductions
    exp = torch.zeros(1000, dtype=torch.float64, device=device)
    for f in st.get_table("fact").fragments:
        fk = f.device_chunks["fk"].view(torch.int32).to(torch.int64)
        m = fk < dim_rows
        exp.scatter_add_(0, fk[m] % 1000, f.device_chunks["measure"].view(torch.float64)[m])
    dist.all_reduce(exp)
    cols, n = _compact(ex, pq, prep)
    order = torch.argsort(cols[0])
    ok = n == 1000 and torch.allclose(cols[1][order].view(torch.float64), exp, rtol=1e-9, atol=0)
    digest = torch.tensor([int(cols[1][order].sum())], dtype=torch.int64, device=device)       # sum of the bit patterns
    lo, hi = digest.clone(), digest.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    same = int(lo.item()) == int(hi.item())
    bad = _sum_over_ranks(0 if (ok and same) else 1, device)
    return {"workload": "BASELINE.json configs[4], weak scaling: star join + group-by SUM, fact rows sharded, dimension broadcast",
            "n_gpus": world, "rows_per_gpu": rows_per_gpu, "ms": ms, "rows_per_s": rows_per_gpu * world / (ms * 1e-3),
            "achieved_gbs_per_gpu": rows_per_gpu * benchdata.C5_BYTES_PER_ROW / (ms * 1e-3) / 1e9, "merge": "p2p" if xchg is not None else "nccl",
            "join_build": build, "broadcast_bytes": 10_000_000 * 4 * 2 + 10_000_000 // 8,
            "parity_check": "ok: merged SUM per group within 1e-9 of the all-reduced torch reductions, identical bits on every rank" if bad == 0
            else f"FAILED on {bad} rank(s)"}


def multi_c4(device, rank, world, rows_per_gpu, distinct_total, reps=2):
    """High-cardinality group-by across ranks: rows re-partitioned by key hash — the scatter kernel writes straight into the
    owners' peer buffers (it IS the all-to-all) — then aggregated by their owner with the partitioned aggregation."""
    import torch.distributed as dist
    from hdk_b200 import distributed as D
    st = ArrowStorage()
    benchdata.make_c4(st, device, rows_per_gpu, distinct_total, rank=rank)
    st.get_table("c4").shard = (rank, world)
    ex = Executor(st, device=device.index)
    unit = sql.parse(benchdata.C4_QUERY, st.tables)
    pq = ex.plan(unit, 2 * distinct_total // world + 1024)
    assert pq.qmd.hash_type == abi.BASELINE_HASH
    prep = ex.prepare(pq)
    outer = st.get_table("c4")
    widths = [outer.columns[c].phys_width for c in pq.columns]
    L, stp = ex.lib, ex.ctx.stream_ptr()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    phases, sent_away = [], 0
    prep2 = None
    for i in range(reps + 1):
        dist.barrier()
        torch.cuda.synchronize()
        ev[0].record()
        counts = torch.zeros(world, dtype=torch.int64, device=device)
        _lib.check(L.hdk_b200_shuffle_count(C.byref(pq.plan), C.byref(prep["kp"]), world, counts.data_ptr(), stp), "shuffle_count")
        ev[1].record()
        frag, n_recv = ex._exchange_rows_over_peer_memory(pq, prep, counts, widths)
        ev[2].record()
        del prep2
        prep2 = ex.prepare(pq, fragments=[frag])
        info = ex.launch(pq, prep2)
        ev[3].record()
        torch.cuda.synchronize()
        code = ex._agree_on_error(prep2["err"])
        assert code == 0, f"in-band error {code}"
        if i >= 1:
            phases.append([ev[k].elapsed_time(ev[k + 1]) for k in range(3)])
        c = counts.cpu()
        sent_away = int(c.sum() - c[rank])
    avg = [sum(p[k] for p in phases) / len(phases) for k in range(3)]
    ms = _max_over_ranks(sum(avg), device)
    phase_max = [_max_over_ranks(a, device) for a in avg]
    # parity over all ranks: rows conserved, integer sums bit-exact, checksum of keys x counts, no key on two ranks
    s_v = sum(int(f.device_chunks["v"].view(torch.int64).sum()) for f in outer.fragments)
    s_k1 = 0
    for f in outer.fragments:
        s_k1 = (s_k1 + int(f.device_chunks["k1"].view(torch.int64).sum())) & (2**64 - 1)
    cols, n = _compact(ex, pq, prep2)
    k1, sv, cnt = cols[0], cols[2], cols[3]
    tot = torch.tensor([int(cnt.sum()), int(sv.sum()), s_v, n, int(torch.unique(k1).numel())], dtype=torch.int64, device=device)
    dist.all_reduce(tot)
    ck = torch.tensor([int((k1 * cnt).sum()), s_k1 - (1 << 64) if s_k1 >= (1 << 63) else s_k1], dtype=torch.int64, device=device)
    dist.all_reduce(ck)                      # int64 all-reduce wraps modulo 2^64 like the sums themselves
    gathered = [torch.empty(2, dtype=torch.int64, device=device) for _ in range(world)]
    kmin_max = torch.tensor([int(k1.min()) if n else 0, int(k1.max()) if n else 0], dtype=torch.int64, device=device)
    dist.all_gather(gathered, kmin_max)
    total_rows = rows_per_gpu * world
    problems = []
    if int(tot[0]) != total_rows:
        problems.append(f"sum COUNT {int(tot[0])} != {total_rows}")
    if int(tot[1]) != int(tot[2]):
        problems.append("sum SUM(v)")
    if int(ck[0]) != int(ck[1]):
        problems.append("sum k1 x COUNT")
    if int(tot[3]) != int(tot[4]) or int(tot[3]) > distinct_total:
        problems.append(f"{int(tot[3])} groups, {int(tot[4])} distinct k1")
    sent_total = _sum_over_ranks(sent_away, device)
    return {"workload": "BASELINE.json configs[3], weak scaling: composite-key group-by, key-hash shuffle over NVLink + local aggregate",
            "n_gpus": world, "rows_per_gpu": rows_per_gpu, "distinct_total": distinct_total, "groups": int(tot[3]), "ms": ms,
            "rows_per_s": total_rows / (ms * 1e-3),
            "phases_ms_max_over_ranks": {"count": phase_max[0], "scatter_into_peer_buffers (incl. count all-gather, barriers)": phase_max[1],
                                         "local_aggregate (init + partitioned aggregation)": phase_max[2]},
            "limiting_phase": ["count", "scatter + exchange", "local aggregate"][int(np.argmax(phase_max))],
            "nvlink_bytes": sent_total * benchdata.C4_BYTES_PER_ROW, "nvlink_gbs_per_gpu": (sent_total / world) * benchdata.C4_BYTES_PER_ROW / (phase_max[1] * 1e-3) / 1e9,
            "local_aggregate_path": "radix-partitioned" if int(info.strategy) == abi.STRATEGY_PARTITIONED else "global-table probe",
            "parity_check": "ok: rows conserved, sum SUM(v) bit-exact, sum k1 x COUNT (mod 2^64) equal over all ranks, every key on exactly one rank"
            if not problems else "FAILED: " + "; ".join(problems)}
