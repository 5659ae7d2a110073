#!/usr/bin/env python
"""Digest of the round-2 `ncu --set full` captures of the other named configs (tools/prof_r2.sh exports the reports on
the GPU box with `ncu -i … --page raw --csv`; the reports themselves are too big to pull):
  profiles/r2_other_configs.csv    selected counters of the second (warm) launch of every kernel
  profiles/traffic.json            + dram bytes per launch for c1_int64, c1_fp64, tpch_q1, c5_star_join, c4_baseline_hash
                                   (benchcfg.py → per_config[*].traffic)

    python tools/profile_configs.py gpurun_out/r2_prof_cfg_raw.csv gpurun_out/r2_prof_c4_raw.csv
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__shared_mem_per_block_dynamic", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "lts__t_sector_hit_rate.pct", "smsp__inst_executed_op_shared_atom.sum",
]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def load(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in rows[2:]:
        d = {"name": r[idx["Kernel Name"]]}
        for m in METRICS:
            if m in idx:
                d[m] = (r[idx[m]].replace(",", ""), units[idx[m]])
        out.append(d)
    return out


def dram(d):
    return sum(float(d[k][0]) * UNIT[d[k][1]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))


def main():
    cfg, c4 = load(sys.argv[1]), load(sys.argv[2])
    # launch order of tools/bench_configs.py --only c1,tpch,c5 --iters 2 (scan kernels only) and --only c4 --iters 2 (pa_* only)
    labels = ["c1_int64", "c1_int64", "c1_fp64", "c1_fp64", "tpch_q1", "tpch_q1", "tpch_q6", "tpch_q6", "c5_star_join", "c5_star_join"]
    picked = [(labels[i], cfg[i]) for i in (1, 3, 5, 7, 9)]
    c4_second = c4[5:10]
    assert [d["name"].split("<")[0].split("(")[0] for d in c4_second][:1] == ["void pa_count_kernel"] or "pa_count" in c4_second[0]["name"]
    picked += [("c4_" + d["name"].replace("void ", "").split("(")[0].split("<")[0] + ("_L" + d["name"].split(", ")[1][0] if "scatter" in d["name"] else ""), d)
               for d in c4_second]
    with open(os.path.join(ROOT, "profiles", "r2_other_configs.csv"), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["config / kernel", "kernel name"] + METRICS)
        for label, d in picked:
            w.writerow([label, d["name"]] + [" ".join(d.get(m, ("", ""))) for m in METRICS])
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    t = json.load(open(tpath))
    for label, d in picked[:5]:
        if label != "tpch_q6":
            t[label] = dram(d)
    t["c4_baseline_hash"] = sum(dram(d) for d in c4_second)
    t["_source_r2"] = ("r2 keys (c1_*, tpch_q1, c5_star_join, c4_baseline_hash = all five kernels of the partitioned aggregation): "
                       "dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full at BASELINE.json sizes (tools/prof_r2.sh, "
                       "profiles/r2_other_configs.csv)")
    json.dump(t, open(tpath, "w"), indent=1)
    for label, d in picked:
        print(label, d["gpu__time_duration.sum"], round(dram(d) / 1e9, 3), "GB")
    print("c4 total", round(t["c4_baseline_hash"] / 1e9, 2), "GB")


if __name__ == "__main__":
    main()
