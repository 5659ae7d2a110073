#!/usr/bin/env python
"""Weak-scaling timings of the other billion-row configs on N GPUs of one box (one process per GPU, torchrun):
  c5  star join + group-by   : fact rows sharded, dimension replicated, partials merged over peer memory
                               (hdk_b200_launch_exchange)
  c4  baseline-hash group-by : rows re-partitioned by key hash (shuffle_count → shuffle_scatter → NCCL all-to-all),
                               local aggregate; disjoint results
Device time by CUDA events, max over ranks; not the bench contract (bench.py is).

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_multigpu.py --scale 0.25
"""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import benchdata  # noqa: E402
from hdk_b200 import _lib, abi, distributed as D, sql  # noqa: E402
from hdk_b200.executor import Executor  # noqa: E402
from hdk_b200.storage import ArrowStorage, ChunkStats, Fragment  # noqa: E402


def max_over_ranks(ms, device):
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    if D.world() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


def sync_stats(st, name, device):
    """all ranks must plan with the same key ranges"""
    if D.world() == 1:
        return
    tab = st.get_table(name)
    for cname, ci in tab.columns.items():
        lo, hi, hn = tab.col_stats(cname)
        tt = torch.tensor([float(lo), -float(hi)], dtype=torch.float64, device=device)
        dist.all_reduce(tt, op=dist.ReduceOp.MIN)
        for f in tab.fragments:
            f.stats[cname].min = tt[0].item() if ci.type.is_fp else int(tt[0].item())
            f.stats[cname].max = -tt[1].item() if ci.type.is_fp else int(-tt[1].item())


def bench_c5(args, rank, world, device):
    rows = int(2_000_000_000 * args.scale)
    st = ArrowStorage()
    benchdata.make_star(st, device, rows, 10_000_000, rank=rank)
    sync_stats(st, "fact", device)
    ex = Executor(st, device=device.index)
    pq = ex.plan(sql.parse(benchdata.C5_QUERY, st.tables))
    prep = ex.prepare(pq)          # builds the join table + slot-ordered payload on this rank (dimension is replicated)
    prep["scratch"] = torch.empty(prep["scratch_bytes"] + 128, dtype=torch.uint8, device=device)
    xchg = D.PeerExchange(ex.lib, pq.plan, pq.qmd, device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for i in range(args.reps + 2):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0.record()
        ex.launch_exchange(pq, prep, xchg)
        e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            ts.append(e0.elapsed_time(e1))
    assert int(prep["err"].item()) == 0
    ms = max_over_ranks(sum(ts) / len(ts), device)
    xchg.close()
    return {"config": "c5 star join + group-by (weak scaling, merge over peer memory)", "n_gpus": world, "rows_per_gpu": rows,
            "ms": round(ms, 3), "rows_per_s": rows * world / (ms * 1e-3), "gbs_per_gpu": round(rows * benchdata.C5_BYTES_PER_ROW / ms / 1e6, 1)}


def bench_c4(args, rank, world, device):
    rows = int(1_000_000_000 * args.scale)
    distinct = int(100_000_000 * args.scale) * world
    st = ArrowStorage()
    benchdata.make_c4(st, device, rows, distinct, rank=rank)
    sync_stats(st, "c4", device)
    ex = Executor(st, device=device.index)
    L = ex.lib
    unit = sql.parse(benchdata.C4_QUERY, st.tables)
    pq = ex.plan(unit, 2 * distinct // world)
    assert pq.qmd.hash_type == abi.BASELINE_HASH
    prep = ex.prepare(pq)
    outer = st.get_table("c4")
    widths = [outer.columns[c].phys_width for c in pq.columns]
    stp = ex.ctx.stream_ptr()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    phases = []
    for i in range(args.reps + 1):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev[0].record()
        counts = torch.zeros(world, dtype=torch.int64, device=device)
        _lib.check(L.hdk_b200_shuffle_count(C.byref(pq.plan), C.byref(prep["kp"]), world, counts.data_ptr(), stp), "shuffle_count")
        ev[1].record()
        if args.a2a == "p2p":
            frag, n_recv = ex._exchange_rows_over_peer_memory(pq, prep, counts, widths)
        else:
            n_local = int(counts.sum().item())
            offsets = torch.cumsum(counts, 0) - counts
            cursors = torch.zeros(world, dtype=torch.int64, device=device)
            send_cols = [torch.empty(max(n_local, 1) * w, dtype=torch.uint8, device=device) for w in widths]
            ptrs = torch.tensor([t.data_ptr() for t in send_cols], dtype=torch.int64, device=device)
            _lib.check(L.hdk_b200_shuffle_scatter(C.byref(pq.plan), C.byref(prep["kp"]), world, offsets.data_ptr(), cursors.data_ptr(),
                                                  ptrs.data_ptr(), stp), "shuffle_scatter")
            recv_cols, n_recv = D.all_to_all_rows(send_cols, counts, widths)
            frag = Fragment(0, n_recv, 0, 0, {}, {c: ChunkStats(None, None, False) for c in pq.columns},
                            {c: t[: n_recv * w] for c, t, w in zip(pq.columns, recv_cols, widths)})
        ev[2].record()
        prep2 = ex.prepare(pq, fragments=[frag])
        ex.launch(pq, prep2)
        ev[3].record()
        torch.cuda.synchronize()
        code = int(prep2["err"].item())
        assert code == 0, f"in-band error {code}"
        if i >= 1:
            phases.append([ev[k].elapsed_time(ev[k + 1]) for k in range(3)])
        del prep2, frag
    avg = [sum(p[k] for p in phases) / len(phases) for k in range(3)]
    ms = max_over_ranks(sum(avg), device)
    return {"config": "c4 baseline-hash group-by (weak scaling, key-hash shuffle + all-to-all)", "n_gpus": world, "rows_per_gpu": rows,
            "distinct_total": distinct, "ms": round(ms, 3), "rows_per_s": rows * world / (ms * 1e-3),
            "exchange": args.a2a,
            "phases_ms_rank0": {"count": round(avg[0], 2), "scatter + exchange (incl. count all-gather, barriers)": round(avg[1], 2),
                                "aggregate(init+scan+finalize)": round(avg[2], 2)}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=0.25)
    ap.add_argument("--only", default="c5,c4")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--a2a", choices=["p2p", "nccl"], default="p2p", help="c4: scatter into peer memory (fused) or scatter + NCCL all-to-all")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    for name in args.only.split(","):
        res = {"c5": bench_c5, "c4": bench_c4}[name](args, rank, world, device)
        if rank == 0:
            print(json.dumps(res), flush=True)
        torch.cuda.empty_cache()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
