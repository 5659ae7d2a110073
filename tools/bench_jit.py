#!/usr/bin/env python
"""A query that is NOT one of the pre-compiled shapes (taxi Q2 with an extra MAX) at the benchmark's size: the interpreting
kernel against the kernels NVRTC builds for the shape at run time, and the compile latency of the first query.

    python tools/bench_jit.py [--rows 1100000000]
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import benchdata  # noqa: E402
from hdk_b200 import _lib, abi, sql  # noqa: E402
from hdk_b200.executor import Executor  # noqa: E402
from hdk_b200.storage import ArrowStorage  # noqa: E402

QUERIES = {
    "taxi_q2_plus_max": ("SELECT passenger_count, avg(total_amount), max(trip_distance) FROM trips GROUP BY passenger_count", 18),
    "taxi_q3_filtered": ("SELECT passenger_count, extract(year from pickup_datetime) AS y, count(*), sum(total_amount) FROM trips "
                         "WHERE trip_distance > 1.5 GROUP BY passenger_count, y", 26),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=1_100_000_000)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    st = ArrowStorage()
    benchdata.make_taxi(st, dev, args.rows)
    ex = Executor(st)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for name, (text, bytes_per_row) in QUERIES.items():
        pq = ex.plan(sql.parse(text, st.tables))
        prep = ex.prepare(pq)
        out = {"query": name, "sql": text, "rows": args.rows, "bytes_per_row": bytes_per_row}
        for mode, jit in (("interpreted", 0), ("specialised", 2)):
            _lib.debug_set("jit", jit)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            info = ex.launch(pq, prep)
            torch.cuda.synchronize()
            first_ms = (time.perf_counter() - t0) * 1e3
            ts = []
            for _ in range(5):
                e0.record()
                info = ex.launch(pq, prep)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ms = sum(ts) / len(ts)
            gbs = bytes_per_row * args.rows / (ms * 1e-3) / 1e9
            out[mode] = {"launch_ms": round(ms, 3), "gbs": round(gbs, 1), "frac_of_measured_peak": round(gbs / peak, 3), "variant": int(info.variant),
                         "strategy": int(info.strategy), "first_launch_wall_ms": round(first_ms, 1)}
            assert int(prep["err"].item()) == 0
        out["jit_stats"] = _lib.jit_stats()
        print(json.dumps(out), flush=True)
    _lib.debug_set("jit", 1)


if __name__ == "__main__":
    main()
