#!/usr/bin/env python
"""Dynamic SASS opcode mix per CUDA source line from an `ncu --page source --csv --print-source cuda,sass`
dump (source rows are followed by the SASS rows correlated with them).
Usage: python profiles/ncu_opmix.py both.csv <first_line> <last_line> [num_envs] [file]"""
import csv
import sys
from collections import Counter, defaultdict


def main(path, lo, hi, nenv=65536.0, fname="ssd_step_fast.cu"):
    cur_file, cur_line, hdr = None, None, None
    byline, src, total = defaultdict(Counter), {}, Counter()
    for r in csv.reader(open(path)):
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = r
        elif hdr and len(r) > 10 and r[0] != "Function Name":
            if r[0]:
                cur_line = int(r[0])
                src[(cur_file, cur_line)] = r[1].strip()[:80]
            elif r[2].startswith("0x"):
                t = r[3].split()
                try:
                    ex = int(r[7])
                except ValueError:
                    continue
                op = (t[1] if t[0].startswith("@") and len(t) > 1 else t[0]).split(".")[0]
                total[op] += ex
                if cur_file == fname:
                    byline[cur_line][op] += ex
    print("whole kernel:", {k: round(v / nenv, 1) for k, v in total.most_common(14)})
    for ln in sorted(byline):
        if lo <= ln <= hi:
            tot = sum(byline[ln].values())
            if tot / nenv > 0.8:
                print(ln, round(tot / nenv, 1), {k: round(v / nenv, 1) for k, v in byline[ln].most_common(7)}, "|", src[(fname, ln)])


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4]) if len(sys.argv) > 4 else 65536.0,
         sys.argv[5] if len(sys.argv) > 5 else "ssd_step_fast.cu")
