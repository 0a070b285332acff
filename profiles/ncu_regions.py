#!/usr/bin/env python
"""Per-function-region and per-stall-reason totals of an `ncu --page source --csv --print-source cuda,sass`
dump of the step kernels.  Regions are found from the `__device__`/`__global__` function headers of
csrc/ssd_phases.cuh, ssd_step_fast.cu and ssd_step_general.cu, so the script follows the source as it changes.
(ncu attributes an inlined instruction to the call site AND the callee line: the per-region sums overlap a little.)
Usage: python profiles/ncu_regions.py both.csv <num_envs_per_launch> [num_launches_in_dump]"""
import csv
import os
import re
import sys

CSRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "sequential_social_dilemma_games_b200", "csrc")
FILES = ("ssd_phases.cuh", "ssd_step_fast.cu", "ssd_step_general.cu")  # older dumps: one file, ssd_step.cu


def regions(fname):
    out, cur = [], None
    path = os.path.join(CSRC, fname)
    if not os.path.exists(path):
        return out
    for ln, line in enumerate(open(path), 1):
        m = re.match(r"^(?:template.*\n)?(?:static\s+)?(?:__device__|__global__)[^(]*?\b(\w+)\s*\(", line)
        if m:
            cur = m.group(1)
            out.append([cur, ln, 10 ** 9])
            if len(out) > 1:
                out[-2][2] = ln - 1
        m2 = re.match(r"^\s*// ---- (.*)$", line)
        if m2 and cur in ("ssd_step_kernel", "ssd_step_fast_kernel"):
            out.append(["kernel: " + m2.group(1)[:40], ln, 10 ** 9])
            out[-2][2] = ln - 1
    return out


def main(path, n_envs, n_launch=1):
    cur_file, hdr = None, None
    inst, stall, stalls_by = {}, {}, {}
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
                key = (cur_file, int(r[0]))
                inst[key] = inst.get(key, 0) + int(d["Instructions Executed"])
                stall[key] = stall.get(key, 0) + int(d["Warp Stall Sampling (All Samples)"])
            except (ValueError, KeyError):
                continue
            for k, v in d.items():
                if k.startswith("stall_") and "Not Issued" not in k:
                    try:
                        stalls_by[k] = stalls_by.get(k, 0) + int(v)
                    except ValueError:
                        pass
    ti, ts = sum(inst.values()), sum(stall.values())
    print("warp instructions per env-step: %.1f   (total %d over %d launch(es) of %d envs)" % (ti / n_launch / n_envs, ti, n_launch, n_envs))
    print("%-44s %7s %7s %12s" % ("region", "inst%", "stall%", "inst/env"))
    acc_i = acc_s = 0
    for fname in FILES:
        for name, a, b in regions(fname):
            i = sum(v for (f, l), v in inst.items() if f == fname and a <= l <= b)
            s = sum(v for (f, l), v in stall.items() if f == fname and a <= l <= b)
            acc_i += i
            acc_s += s
            if i:
                print("%-44s %6.2f%% %6.2f%% %12.1f" % (name, 100.0 * i / ti, 100.0 * s / max(ts, 1), i / n_launch / n_envs))
    print("%-44s %6.2f%% %6.2f%% %12.1f" % ("other files (philox, intrinsics, atomics)", 100.0 * (ti - acc_i) / ti, 100.0 * (ts - acc_s) / max(ts, 1), (ti - acc_i) / n_launch / n_envs))
    print()
    tot = sum(stalls_by.values())
    for k, v in sorted(stalls_by.items(), key=lambda x: -x[1])[:10]:
        print("%-28s %6.2f%%" % (k, 100.0 * v / max(tot, 1)))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 1)
