#!/usr/bin/env python
"""Golden fixture of the catalog generator's acceptance rule, from the REFERENCE's own functions (run in the build
container; needs /root/reference):  fx = envs/farnocchia.py::fx_xyz_farnocchia (numba, unmodified), ecef2lla /
ecef2aer = envs/transformations.py (astropy._erfa stub of oracle/ref_loader.py), hx_aer_erfa = its 5-line body
(dynamics.py:219-231, oracle/dynamics_restated.py), the rule = orbit_gen.py:60-73 transcribed in
oracle/orbit_gen_oracle.py::accept_rule.  -> tests/golden/golden_orbit_gen.npz"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from datetime import datetime  # noqa: E402

from oracle import dynamics_restated as D  # noqa: E402
from oracle import orbit_gen_oracle as OG  # noqa: E402
from oracle import ref_loader as rl  # noqa: E402
from ssa_gym_b200.orbit_gen import sample_candidates  # noqa: E402
from ssa_gym_b200.transformations import gcrs2irts_matrix_approx, time_table  # noqa: E402


def main():
    far, tr = rl.farnocchia(), rl.transformations()
    fx = far.fx_xyz_farnocchia
    hx, _, _, _ = D.make_operators(tr)
    step, n = 150.0, 96
    table = gcrs2irts_matrix_approx(time_table(datetime(2020, 5, 4), step, n))
    lla = np.array([np.radians(38.828198), np.radians(-77.305352), 20.0])
    obs_itrs = tr.lla2ecef(lla)
    cand = sample_candidates(400, np.random.RandomState(7))
    K = len(cand)
    el = np.zeros((K, n)); alt = np.zeros((K, n))
    for c in range(K):
        for i in range(n):
            x = fx(cand[c], step * i)
            el[c, i] = hx(x, table[i], lla, obs_itrs)[1]
            alt[c, i] = tr.ecef2lla(x[:3] @ table[i])[2]
    acc = np.array([OG.accept_rule(alt[c], el[c], np.radians(15), step) for c in range(K)])
    print("accepted", acc.sum(), "of", K)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_orbit_gen.npz"), candidates=cand,
                        table=table, lla=lla, obs_itrs=obs_itrs, step=step, elevation=el, altitude=alt, accept=acc)


if __name__ == "__main__":
    main()
