#!/usr/bin/env python
"""Instructions and stall samples of the update body per PHASE (hx, mean_z, residuals, S / Pxz, inverse, gain, write-back,
epilogue): every SASS instruction of an `ncu --page source --print-source cuda,sass --csv` export is attributed, in
address order, to the phase of the most recent instruction that maps to a line of tile_update_body — so the inlined
math (ssa_math.h, ssa_meas.h) lands in the phase that called it.  python tools/phase_mix.py export.csv"""
import collections
import csv
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MARKS = [("hx phase", "// ---- 14 T measurements: task = (object, sigma index); 13 = the TRUE state (visibility, z_true)"),
         ("mean_z", "// ---- mean_z: Cartesian (uvw) mean"), ("residuals", "// ---- the 13 residuals"),
         ("S / Pxz", "// ---- cross covariance (thread per state row"), ("inverse", "// ---- S^-1 (numpy.linalg.inv"),
         ("gain", "// ---- gain row, S K^T column"), ("write-back", "// ---- P -= K (S K^T) and the new mean go back"),
         ("epilogue", "// ---- epilogue (results.py:36-72)"), ("staging (wrapper)", "template <int T, int NT, bool REFACTOR>")]


def main(path):
    src = open(os.path.join(ROOT, "ssa_gym_b200", "csrc", "ssa_tile.cuh")).read().splitlines()
    marks = sorted((next(i for i, l in enumerate(src, 1) if m in l), n) for n, m in MARKS)
    body0 = marks[0][0] - 25

    def phase(line):
        name = None
        for l, n in marks:
            if line >= l:
                name = n
        return name

    recs, cur, line, hdr = [], None, None, None
    for r in csv.reader(open(path)):
        if len(r) == 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r and r[0] == "Line No":
            hdr = r
            iA, iE, iS = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
        elif hdr and len(r) > iE:
            if r[0].isdigit():
                line = int(r[0])
            elif r[iA].startswith("0x") and r[iE].isdigit():
                recs.append((int(r[iA], 16), cur, line, int(r[iE]), int(r[iS]) if r[iS].isdigit() else 0))
    recs.sort()
    ph = "staging (wrapper)"
    pi, ps = collections.Counter(), collections.Counter()
    for _, f, l, e, s in recs:
        if f == "ssa_tile.cuh" and l >= body0 and phase(l):
            ph = phase(l)
        pi[ph] += e
        ps[ph] += s
    ti, ts = sum(pi.values()), sum(ps.values())
    for k in pi:
        print(f"{k:20s} instructions {100 * pi[k] / ti:5.1f} %   stall samples {100 * ps[k] / ts:5.1f} %")


if __name__ == "__main__":
    main(sys.argv[1])
