"""Load the importable pieces of the read-only reference tree by path (TEST INFRASTRUCTURE).

Only `tests/golden/make_golden.py` (and ad-hoc validation here in the build container) uses this:
`/root/reference` does not exist on the GPU box, so nothing that runs there may depend on it.

What can be imported in this image:
  * envs/farnocchia.py          — imports unmodified (numpy + numba)
  * envs/transformations.py     — needs `astropy._erfa`; a stub providing the five constants/functions
                                   the module touches at import time is injected.  The njit geometry
                                   (lla2ecef, ecef2aer, aer2uvw, uvw2aer, ecef2lla) then works; the ERFA
                                   matrix builders do not (ERFA is absent).
filterpy, gym, astropy, poliastro are absent, so envs/__init__.py, dynamics.py and
ssa_tasker_simple_2.py cannot be imported; their few path functions are restated in oracle/.
"""
import importlib.util
import os
import sys
import types
import warnings

REF_ROOT = os.environ.get("SSA_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "envs", "farnocchia.py"))


def _load(name, relpath):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF_ROOT, relpath))
    mod = importlib.util.module_from_spec(spec)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        spec.loader.exec_module(mod)
    return mod


_cache = {}


def farnocchia():
    if "far" not in _cache:
        _cache["far"] = _load("_ref_farnocchia", "envs/farnocchia.py")
    return _cache["far"]


def transformations():
    if "tr" not in _cache:
        if "astropy" not in sys.modules:
            astropy = types.ModuleType("astropy")
            erfa = types.ModuleType("astropy._erfa")
            erfa.DAYSEC = 86400.0
            erfa.DAS2R = 4.848136811095359935899141e-6
            erfa.DMAS2R = erfa.DAS2R / 1e3
            erfa.DPI = 3.141592653589793238462643
            erfa.eform = lambda n: (6378137.0, 1.0 / 298.257223563)  # WGS84
            astropy._erfa = erfa
            sys.modules["astropy"] = astropy
            sys.modules["astropy._erfa"] = erfa
        _cache["tr"] = _load("_ref_transformations", "envs/transformations.py")
    return _cache["tr"]


def catalog():
    import numpy as np
    return np.load(os.path.join(REF_ROOT, "envs", "1.5_hour_viz_20000_of_20000_sample_orbits_seed_0.npy"))
