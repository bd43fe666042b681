#!/usr/bin/env python
"""Per source line executed warp instructions and stall samples of one kernel from an
`ncu --page source --print-source cuda,sass --csv` export.  python tools/src_hot.py export.csv [top]"""
import collections
import csv
import sys


def main(path, top=60):
    cur, hdr = None, None
    inst, samp, text = collections.Counter(), collections.Counter(), {}
    for r in csv.reader(open(path)):
        if len(r) == 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if r and r[0] == "Line No":
            hdr = r
            iL, iSrc, iE, iS = 0, 1, hdr.index("Instructions Executed"), hdr.index("# Samples")
            continue
        if hdr is None or len(r) <= iE or not r[iL].isdigit():
            continue
        if r[iE].isdigit():
            key = (cur, int(r[iL]))
            inst[key] += int(r[iE])
            samp[key] += int(r[iS]) if r[iS].isdigit() else 0
            if r[iSrc].strip():
                text[key] = r[iSrc].strip()
    tot, tots = sum(inst.values()), sum(samp.values())
    byfile = collections.Counter()
    sbyfile = collections.Counter()
    for (f, l), n in inst.items():
        byfile[f] += n
        sbyfile[f] += samp[(f, l)]
    print("total", tot, "samples", tots)
    for f, n in byfile.most_common():
        print(f"  {f:20s} {100 * n / tot:5.1f}% inst  {100 * sbyfile[f] / max(1, tots):5.1f}% samples")
    for key, n in inst.most_common(top):
        print(f"{key[0]:16s}:{key[1]:5d} {100 * n / tot:5.2f}% inst {100 * samp[key] / max(1, tots):5.2f}% smp  {text.get(key, '')[:110]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 60)
