#!/usr/bin/env python
"""Hot source lines of one kernel in an `ncu --set full --import-source on` report (built with -lineinfo):
instructions executed and stall samples per source line.   python tools/ncu_lines.py rep kernel-regex [top]"""
import csv
import io
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    sel = ["--kernel-id", kern] if kern.startswith(":") else ["--kernel-name", f"regex:{kern}"]   # ":::3" = third launch
    raw = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"] + sel,
                                  text=True, stderr=subprocess.DEVNULL)
    fname, hdr, data = "", None, []
    for r in csv.reader(io.StringIO(raw)):
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = r
        elif hdr and r[0].isdigit() and r[2] == "-":      # a source line (its SASS lines follow with an address)
            try:
                data.append((int(r[hdr.index("Instructions Executed")]), int(r[hdr.index("# Samples")]), fname, r[0], r[1].strip()[:140]))
            except ValueError:
                pass
    ti, ts = sum(d[0] for d in data), sum(d[1] for d in data)
    print(f"total warp instructions {ti}, samples {ts}")
    for d in sorted(data, reverse=True)[:top]:
        print(f"{100 * d[0] / max(ti, 1):5.1f}% inst {100 * d[1] / max(ts, 1):5.1f}% smp  {d[2]}:{d[3]}  {d[4]}")


if __name__ == "__main__":
    main()
