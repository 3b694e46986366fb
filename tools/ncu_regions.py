"""Per-region instruction / stall-sample shares of one kernel from an `ncu --page source --csv` export.
usage: python tools/ncu_regions.py <source.csv> <regions.txt>
regions.txt: lines `name file first last` (source file basename, inclusive line range); unmatched lines go to `other`."""
import csv
import sys
from collections import defaultdict


def rows(path):
    fn = None
    for r in csv.reader(open(path)):
        if not r:
            continue
        if r[0] == "File Path":
            fn = r[1].split("/")[-1]
            continue
        if r[0].isdigit():
            def I(x):
                try:
                    return int(x)
                except ValueError:
                    return 0
            yield fn, int(r[0]), I(r[7]), I(r[4])


def main():
    regions = []
    for ln in open(sys.argv[2]):
        p = ln.split()
        if len(p) == 4:
            regions.append((p[0], p[1], int(p[2]), int(p[3])))
    inst, stall = defaultdict(int), defaultdict(int)
    for fn, line, ni, ns in rows(sys.argv[1]):
        name = "other"
        for rn, rf, a, b in regions:
            if fn == rf and a <= line <= b:
                name = rn
                break
        inst[name] += ni
        stall[name] += ns
    ti, ts = sum(inst.values()), sum(stall.values())
    print(f"total warp instructions {ti}, stall samples {ts}")
    for k in sorted(inst, key=lambda k: -inst[k]):
        print(f"  {k:28s} {100 * inst[k] / ti:5.1f}% inst  {100 * stall[k] / max(ts, 1):5.1f}% stall samples")


if __name__ == "__main__":
    main()
