#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line:
executed warp instructions, stall samples, shared-memory wavefronts.  Usage:
    ncu -i prof.ncu-rep --page source --csv --print-source cuda,sass > both.csv
    python profiles/ncu_by_line.py both.csv [top_n]
"""
import csv
import sys


def main(path, top=45):
    cur_file, hdr, rows = None, None, []
    for r in csv.reader(open(path)):
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = r
        elif hdr and r[0] not in ("", "Function Name") and len(r) > 10:
            d = dict(zip(hdr[4:], r[4:]))
            try:
                rows.append((cur_file, int(r[0]), r[1].strip()[:90], int(d["Instructions Executed"]),
                             int(d["Warp Stall Sampling (All Samples)"]), int(d.get("L1 Wavefronts Shared", "0") or 0),
                             float(d.get("Avg. Threads Executed", "0") or 0)))
            except (ValueError, KeyError):
                pass
    tot_i = sum(r[3] for r in rows)
    tot_s = sum(r[4] for r in rows)
    print("total warp instructions %d, stall samples %d" % (tot_i, tot_s))
    print("%-16s %5s %7s %7s %9s %5s  %s" % ("file", "line", "inst%", "stall%", "smem_wf", "thr", "source"))
    for f, ln, src, ins, st, wf, thr in sorted(rows, key=lambda r: -r[3])[:top]:
        print("%-16s %5d %6.2f%% %6.2f%% %9d %5.1f  %s" % (f, ln, 100.0 * ins / tot_i, 100.0 * st / max(tot_s, 1), wf, thr, src))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 45)
