#!/usr/bin/env python
"""Instruction mix of one kernel from an `ncu --page source --csv` export: executed warp instructions and stall
samples per SASS opcode, and the hottest source regions.  Development tool: python tools/sass_mix.py export.csv"""
import collections
import csv
import re
import sys


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
    hdr, data = rows[h], rows[h + 1:]
    iS, iE, iSamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    tot = 0
    byop, samp = collections.Counter(), collections.Counter()
    for r in data:
        if len(r) <= iE or not r[iE].isdigit():
            continue
        n, s = int(r[iE]), int(r[iSamp]) if r[iSamp].isdigit() else 0
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[iS].strip())
        op = (m.group(2) if m else r[iS].strip()).split(".")[0]
        byop[op] += n
        samp[op] += s
        tot += n
    print("total warp instructions", tot, " samples", sum(samp.values()))
    for op, n in byop.most_common(top):
        print(f"{op:12s} {n:12d} {100 * n / tot:5.1f}%  samples {samp[op]:7d} {100 * samp[op] / max(1, sum(samp.values())):5.1f}%")


if __name__ == "__main__":
    main(sys.argv[1])
