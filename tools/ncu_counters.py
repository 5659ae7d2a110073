#!/usr/bin/env python
"""Print selected counters + top stall reasons of every kernel in ncu reports (read here without a GPU):
    python tools/ncu_counters.py gpurun_out/prof_*.ncu-rep"""
import csv, io, subprocess, sys
METRICS = ["gpu__time_duration.sum","dram__bytes_read.sum","dram__bytes_write.sum","gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
"smsp__issue_active.avg.pct_of_peak_sustained_active","sm__warps_active.avg.pct_of_peak_sustained_active","launch__registers_per_thread","launch__block_size","launch__grid_size",
"launch__shared_mem_per_block_dynamic","smsp__inst_executed.sum","l1tex__throughput.avg.pct_of_peak_sustained_active","lts__throughput.avg.pct_of_peak_sustained_elapsed",
"lts__t_sector_hit_rate.pct","lts__t_sectors.sum","lts__t_sectors_op_read.sum","lts__t_sectors_op_atom.sum","lts__t_sectors_op_red.sum","l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum","l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
"l1tex__data_pipe_lsu_wavefronts_mem_shared.sum","l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum","sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active","sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
"smsp__inst_executed_op_shared_atom.sum","smsp__inst_executed_op_global_ld.sum", "smsp__inst_executed_op_global_atom.sum","smsp__inst_executed_op_global_red.sum"]
for rep in sys.argv[1:]:
    raw = subprocess.check_output(["ncu","-i",rep,"--page","raw","--csv"],text=True)
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h:i for i,h in enumerate(hdr)}
    stall = [h for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio")]
    print("=====", rep)
    for r in data:
        print(r[idx["Kernel Name"]][:90])
        for m in METRICS:
            if m in idx: print(f"  {m:75s} {r[idx[m]]:>18s} {units[idx[m]]}")
        st = sorted(((float(r[idx[h]]),h) for h in stall), reverse=True)[:7]
        for v,h in st: print(f"  stall {h.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''):30s} {v:.2f}")
