#!/usr/bin/env python
"""Measured operation count of the reference's literal per-step sequence (SURVEY.md 8(d) asks for the estimate of
39 kflop per unit to be replaced by a measured one).  Runs the operation-counting build of the CPU oracle
(oracle/opcount.cpp: oracle/ukf_oracle.c with `double` replaced by a counting wrapper — same results, every arithmetic
operation and library call counted) on the C2 inputs and prices the counts with SURVEY 8(d)'s convention.  CPU only.

  python tools/opcount.py [--objects 20000] [--steps 3]     -> prints the table, writes profiles/opcount_reference_sequence.json
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SLOTS = ["add", "mul", "div", "cmp", "sqrt", "sin", "cos", "tan", "atan", "atan2", "asin", "acos", "mod", "hyperbolic", "exp_log", "abs_floor"]
# add / sub / mul / compare-select = 1, div = sqrt = 10, sin = cos = 40, tan = 70, atan = 60, atan2 = 80,
# asin = acos = 70, Python-mod = 10; hyperbolic / exp / log (only in the e >= 1 regimes) priced like tan; abs / floor free
WEIGHT = {"add": 1, "mul": 1, "div": 10, "cmp": 1, "sqrt": 10, "sin": 40, "cos": 40, "tan": 70, "atan": 60, "atan2": 80, "asin": 70,
          "acos": 70, "mod": 10, "hyperbolic": 70, "exp_log": 70, "abs_floor": 0}


def build():
    so = os.path.join(ROOT, "oracle", "liboracle_opcount.so")
    src = os.path.join(ROOT, "oracle", "opcount.cpp")
    dep = os.path.join(ROOT, "oracle", "ukf_oracle.c")
    if not os.path.isfile(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(dep)):
        subprocess.run(["g++", "-O1", "-std=c++17", "-fpermissive", "-w", "-shared", "-fPIC", "-o", so, src], check=True,
                       cwd=os.path.join(ROOT, "oracle"))
    return ctypes.CDLL(so)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--objects", type=int, default=20000)
    ap.add_argument("--steps", type=int, default=3)
    a = ap.parse_args()
    import helpers as H
    from ssa_gym_b200 import _lib as F
    L = build()
    n = a.objects
    cat, x, P0, zn = H.c2_inputs(n, a.steps)
    cfg = H.make_cfg(n)
    p = H.p

    def step(lib, st, flags, s):
        lib.oracle_step(ctypes.byref(cfg), p(np.ascontiguousarray(H.CEL2TER06AXY)), ctypes.c_int(flags), p(st.x_true), p(st.x), p(st.P),
                        p(st.status), p(st.infl), None, p(np.ascontiguousarray(zn[s])), p(st.obs), p(st.dpos), p(st.dvel), p(st.spos),
                        p(st.svel), p(st.trace), p(st.z_true), p(st.y), p(st.S), p(st.sigmas_h), p(st.visible), p(st.updated))

    def counts():
        buf = (ctypes.c_uint64 * L.opcount_slots())()
        L.opcount_get(buf)
        return dict(zip(SLOTS, [int(v) for v in buf]))

    full = F.STEP_TRUTH | F.STEP_PREDICT | F.STEP_UPDATE_ALL | F.STEP_EPILOGUE
    parts = {"unit (truth + predict + update + epilogue)": full, "truth + predict": F.STEP_TRUTH | F.STEP_PREDICT,
             "update + epilogue": F.STEP_UPDATE_ALL | F.STEP_EPILOGUE}
    out = {"objects": n, "steps": a.steps, "convention": WEIGHT, "what": __doc__.split("\n\n")[0]}
    # the counting build must be the same arithmetic as liboracle.so
    st_c, st_o = H.HostState(cat, x, P0), H.HostState(cat, x, P0)
    for s in range(a.steps):
        step(L, st_c, full, s)
        step(H.oracle(), st_o, full, s)
    assert H.bits_equal(st_c.x, st_o.x) and H.bits_equal(st_c.P, st_o.P) and H.bits_equal(st_c.obs, st_o.obs), "counting build differs"
    for name, flags in parts.items():
        st = H.HostState(cat, x, P0)
        tot = None
        for s in range(a.steps):
            if flags != full:  # bring the state to step s with full steps, count the part on top of it
                pass
            L.opcount_reset()
            if flags == full:
                step(L, st, flags, s)
                c = counts()
            else:
                import copy
                st2 = copy.deepcopy(st)
                if flags & F.STEP_UPDATE_ALL and not (flags & F.STEP_PREDICT):  # the update of step s follows its predict
                    step(L, st2, F.STEP_TRUTH | F.STEP_PREDICT, s)
                    L.opcount_reset()
                step(L, st2, flags, s)
                c = counts()
                step(L, st, full, s)
            tot = c if tot is None else {k: tot[k] + c[k] for k in c}
        per = {k: v / (n * a.steps) for k, v in tot.items()}
        flop = sum(per[k] * WEIGHT[k] for k in per)
        strict = sum(per[k] for k in per if k != "abs_floor")
        out[name] = {"per_object": per, "flop_survey_convention": flop, "flop_every_op_1": strict}
        print(f"{name}: {flop / 1e3:.2f} kflop per object by SURVEY 8(d)'s convention ({strict / 1e3:.2f} k with every operation = 1)")
        print("   " + "  ".join(f"{k} {per[k]:.1f}" for k in SLOTS if per[k] > 0.05))
    fx = np.ascontiguousarray(cat[:n])
    o6, exc = np.empty_like(fx), np.zeros(n, np.int32)
    L.opcount_reset()
    L.oracle_fx(p(fx), ctypes.c_double(20.0), p(o6), p(exc), ctypes.c_int(n))
    c = counts()
    per = {k: v / n for k, v in c.items()}
    out["fx (one propagation, dt = 20 s)"] = {"per_call": per, "flop_survey_convention": sum(per[k] * WEIGHT[k] for k in per)}
    print(f"fx: {out['fx (one propagation, dt = 20 s)']['flop_survey_convention'] / 1e3:.3f} kflop per propagation")
    print("   " + "  ".join(f"{k} {per[k]:.2f}" for k in SLOTS if per[k] > 0.005))
    with open(os.path.join(ROOT, "profiles", "opcount_reference_sequence.json"), "w") as f:
        json.dump(out, f, indent=1)
    implemented(a, cat, x, P0, zn, cfg)


TSLOTS = ["add", "mul", "fma", "div", "sqrt", "cmp", "misc"]
TWEIGHT = {"add": 1, "mul": 1, "fma": 2, "div": 10, "sqrt": 10, "cmp": 1, "misc": 1}


def build_twin():
    so = os.path.join(ROOT, "oracle", "libtwin_opcount.so")
    src = os.path.join(ROOT, "oracle", "opcount_twin.cpp")
    deps = [src, os.path.join(ROOT, "tests", "twin", "twin.cpp")] + [os.path.join(ROOT, "ssa_gym_b200", "csrc", h) for h in
                                                                      ("ssa_math.h", "ssa_orbit.h", "ssa_meas.h", "ssa_ukf_core.h")]
    if not os.path.isfile(so) or os.path.getmtime(so) < max(os.path.getmtime(d) for d in deps):
        subprocess.run(["g++", "-O1", "-std=c++17", "-fpermissive", "-w", "-ffp-contract=off", "-mfma", "-shared", "-fPIC", "-o", so, src],
                       check=True, cwd=os.path.join(ROOT, "oracle"))
    return ctypes.CDLL(so)


def implemented(a, cat, x, P0, zn, cfg):
    """The same measurement for the algorithm AS IMPLEMENTED: the host twin (the product's own arithmetic headers) built
    with the counting wrapper — primitive operations only (its sin / cos / atan2 / asin are FMA polynomials), priced
    add = mul = compare = 1, fma = 2, div = sqrt = 10.  -> profiles/opcount_implemented.json"""
    import helpers as H
    from ssa_gym_b200 import _lib as F
    L = build_twin()
    p = H.p
    n = len(cat)

    def step(lib, st, flags, s):
        P = H.pack_P(st.P)
        lib.twin_step(ctypes.byref(cfg), p(np.ascontiguousarray(H.CEL2TER06AXY)), ctypes.c_int(flags), p(st.x_true), p(st.x), p(P),
                      p(st.status), p(st.infl), None, p(np.ascontiguousarray(zn[s])), p(st.obs), p(st.dpos), p(st.dvel), p(st.spos),
                      p(st.svel), p(st.trace), p(st.z_true), p(st.y), p(st.S), p(st.sigmas_h), p(st.visible), p(st.updated))
        st.P = H.unpack_P(P)

    def counts():
        buf = (ctypes.c_uint64 * L.opcount_slots())()
        L.opcount_get(buf)
        return dict(zip(TSLOTS, [int(v) for v in buf]))

    full = F.STEP_TRUTH | F.STEP_PREDICT | F.STEP_UPDATE_ALL | F.STEP_EPILOGUE
    st_c, st_t = H.HostState(cat, x, P0), H.HostState(cat, x, P0)
    tot = None
    for s in range(a.steps):
        L.opcount_reset()
        step(L, st_c, full, s)
        c = counts()
        tot = c if tot is None else {k: tot[k] + c[k] for k in c}
        step(H.twin(), st_t, full, s)
    assert H.bits_equal(st_c.x, st_t.x) and H.bits_equal(st_c.P, st_t.P) and H.bits_equal(st_c.obs, st_t.obs), "counting twin differs"
    per = {k: v / (n * a.steps) for k, v in tot.items()}
    flop = sum(per[k] * TWEIGHT[k] for k in per)
    plain = per["add"] + per["mul"] + 2 * per["fma"]
    fx = np.ascontiguousarray(cat[:n])
    o6, exc = np.empty_like(fx), np.zeros(n, np.int32)
    L.opcount_reset()
    L.twin_fx(p(fx), ctypes.c_double(20.0), p(o6), p(exc), ctypes.c_int(n))
    cf = {k: v / n for k, v in counts().items()}
    out = {"objects": n, "steps": a.steps, "convention": TWEIGHT, "what": implemented.__doc__,
           "unit (truth + predict + update + epilogue)": {"per_object": per, "flop_survey_convention": flop,
                                                          "flop_add_mul_2fma_only": plain},
           "fx (one propagation, dt = 20 s)": {"per_call": cf, "flop_survey_convention": sum(cf[k] * TWEIGHT[k] for k in cf)}}
    print(f"implemented: {flop / 1e3:.2f} kflop per object (add = mul = cmp = 1, fma = 2, div = sqrt = 10); "
          f"{plain / 1e3:.2f} k counting add + mul + 2 fma only")
    print("   " + "  ".join(f"{k} {per[k]:.1f}" for k in TSLOTS))
    print(f"implemented fx: {out['fx (one propagation, dt = 20 s)']['flop_survey_convention'] / 1e3:.3f} kflop per propagation")
    print("   " + "  ".join(f"{k} {cf[k]:.2f}" for k in TSLOTS))
    with open(os.path.join(ROOT, "profiles", "opcount_implemented.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
