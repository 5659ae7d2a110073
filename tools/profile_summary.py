#!/usr/bin/env python
"""Turn an `ncu --set full` capture of the taxi scan kernels (read here, without a GPU, via
`ncu -i <rep> --page raw --csv`) into the tracked summaries under profiles/:
  profiles/<tag>_scan_kernels.csv   selected counters per launch (Q1..Q4 in launch order)
  profiles/traffic.json             dram bytes read+written per launch, keyed q1..q4 (bench.py → roofline.traffic)

    python tools/profile_summary.py gpurun_out/prof_r1_full.ncu-rep r1
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "smsp__inst_executed.sum",
    "l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
    "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
]
UNIT_SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}


def main():
    rep, tag = sys.argv[1], sys.argv[2]
    # (a .csv argument is the `ncu -i … --page raw --csv` export made on the GPU box, for reports too big to pull)
    raw = open(rep).read() if rep.endswith(".csv") else subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True)
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio")]
    cols = ["Kernel Name"] + [m for m in METRICS if m in idx] + stall
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    out = os.path.join(ROOT, "profiles", f"{tag}_scan_kernels.csv")
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["query"] + [c.replace("smsp__average_warps_issue_stalled_", "stall_").replace("_per_issue_active.ratio", "") for c in cols])
        w.writerow(["unit"] + [units[idx[c]] for c in cols])
        for i, r in enumerate(data):
            w.writerow([f"q{i + 1}"] + [r[idx[c]][:60] for c in cols])
    traffic = {}
    for i, r in enumerate(data):
        tot = 0.0
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot += float(r[idx[m]].replace(",", "")) * UNIT_SCALE.get(units[idx[m]], 1.0)
        traffic[f"q{i + 1}"] = tot
    traffic["_source"] = f"{os.path.basename(rep)}: dram__bytes_read.sum + dram__bytes_write.sum per launch (ncu --set full, 1.1B rows)"
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    merged = json.load(open(tpath)) if os.path.exists(tpath) else {}     # keep the other configs' keys (tools/profile_configs.py)
    merged.update(traffic)
    json.dump(merged, open(tpath, "w"), indent=1)
    print("wrote", out, "and profiles/traffic.json", traffic)


if __name__ == "__main__":
    main()
