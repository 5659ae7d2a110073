#!/usr/bin/env python
"""Launch-geometry sweep for the taxi scan kernels on one GPU (tuning aid, not the bench contract).
For each query it times the scan kernel with CUDA events under every candidate
HDK_B200_GEO="strategy,consumer_threads,ctas_per_sm,stages,tile_rows" and under the library's own choice.

    python tools/sweep_geo.py [--rows 400000000] [--only q2,q3]
"""
import argparse
import itertools
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import benchdata  # noqa: E402
from hdk_b200 import sql  # noqa: E402
from hdk_b200.executor import Executor  # noqa: E402
from hdk_b200.storage import ArrowStorage  # noqa: E402


def time_scan(ex, pq, prep, reps):
    e1, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts, info = [], None
    for i in range(reps + 2):
        prep["err"].zero_()
        torch.cuda.synchronize()
        e1.record()
        info = ex.launch(pq, prep)
        e2.record()
        torch.cuda.synchronize()
        if i >= 2:
            ts.append(e1.elapsed_time(e2))
    return sum(ts) / len(ts), info, int(prep["err"].item())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=400_000_000)
    ap.add_argument("--only", default="q1,q2,q3,q4")
    ap.add_argument("--reps", type=int, default=4)
    ap.add_argument("--fine", action="store_true", help="finer grid around the library's own strategy")
    ap.add_argument("--geos", default="", help="semicolon-separated explicit geometries; default: a built-in grid")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    from hdk_b200 import _lib
    _lib.debug_set("geo_env_refresh", 1)     # the library reads HDK_B200_GEO once per process otherwise
    st = ArrowStorage()
    benchdata.make_taxi(st, dev, args.rows)
    ex = Executor(st)
    peak = 6537.0
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = float(json.load(open(p))["hbm_gbs"])
    for q in args.only.split(","):
        unit = sql.parse(benchdata.TAXI_QUERIES[q], st.tables)
        pq = ex.plan(unit, None)
        prep = ex.prepare(pq)
        bpr = benchdata.TAXI_BYTES_PER_ROW[q]
        os.environ.pop("HDK_B200_GEO", None)
        ms, info, err = time_scan(ex, pq, prep, args.reps)
        print(f"{q} default: strategy={info.strategy} block={info.block} grid={info.grid} smem={info.smem_bytes} tile_rows={info.tile_rows} "
              f"ms={ms:.3f} frac={bpr * args.rows / ms / 1e6 / peak:.3f} err={err}", flush=True)
        if args.geos:
            cands = [tuple(int(x) for x in g.split(",")) for g in args.geos.split(";")]
        elif args.fine:
            cands = [(info.strategy, n, c, stg, n * 4 * k) for n in (128, 192, 256, 320, 384, 512) for c in (1, 2, 3, 4, 5, 6)
                     for stg in (2, 3, 4) for k in (1, 2, 3, 4, 6, 8)]
        else:
            cands = [(s, n, c, stg, tr) for s in (0, 1) for n in (128, 256, 384, 512) for c in (1, 2, 3, 4) for stg in (3, 4, 6, 8)
                     for tr in (512, 1024, 2048, 4096, 8192)]
        rows_out = []
        for g in cands:
            os.environ["HDK_B200_GEO"] = ",".join(map(str, g))
            try:
                ms, info, err = time_scan(ex, pq, prep, args.reps)
            except Exception as e:  # does not fit
                continue
            rows_out.append((ms, g, info.smem_bytes, err))
        os.environ.pop("HDK_B200_GEO", None)
        rows_out.sort()
        for ms, g, smem, err in rows_out[:12]:
            print(f"  {q} geo={g} smem={smem} ms={ms:.3f} frac={bpr * args.rows / ms / 1e6 / peak:.3f} err={err}", flush=True)
        for ms, g, smem, err in rows_out[-3:]:
            print(f"  {q} (worst) geo={g} smem={smem} ms={ms:.3f}", flush=True)


if __name__ == "__main__":
    main()
