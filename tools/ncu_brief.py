#!/usr/bin/env python
"""Brief per-kernel digest of an `ncu --set full` report (read without a GPU): time, DRAM bytes, pipe / cache
utilisation, top stall reasons.   python tools/ncu_brief.py gpurun_out/x.ncu-rep [kernel-regex]"""
import csv
import io
import re
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum",
        "lts__t_sector_hit_rate.pct", "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active", "smsp__inst_executed_op_shared_atom.sum"]


def main():
    rep = sys.argv[1]
    pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
    # (a .csv argument is the `--page raw --csv` export itself: reports too big to pull from the GPU box are exported there)
    raw = open(rep).read() if rep.endswith(".csv") else subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True)
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        if pat and not pat.search(name):
            continue
        print("====", name[:90], r[idx["Demangled Name"]][:120] if "Demangled Name" in idx else "")
        for w in WANT:
            if w in idx:
                print(f"  {w:70s} {r[idx[w]]:>16s} {units[idx[w]]}")
        st = []
        for h, i in idx.items():
            if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
                try:
                    st.append((float(r[i]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
                except ValueError:
                    pass
        print("  stalls:", ", ".join(f"{n} {v:.2f}" for v, n in sorted(st, reverse=True)[:7]))


if __name__ == "__main__":
    main()
