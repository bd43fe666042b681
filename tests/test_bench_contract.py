"""CPU: the committed bench lines (profiles/r01_bench_*.json, printed by bench.py on a B200) carry every key of the
measurement contract, with consistent values; and the reference arm runs here (it is CPU-only by definition)."""
import json
import os
import subprocess
import sys

import helpers as H

PROFILES = os.path.join(H.ROOT, "profiles")
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "roofline", "e2e", "gpu_launches", "clocks"}


def _line(name):
    return json.loads(open(os.path.join(PROFILES, name)).read().strip().splitlines()[-1])


def test_headline_bench_line_has_the_contract_keys():
    d = _line("r01_bench_c2.json")
    assert BASE_KEYS | {"cpu_baseline"} <= set(d), BASE_KEYS - set(d)
    assert d["unit"] == "object-updates/s" and d["higher_is_better"] is True and d["scaling"] == "weak" and d["dtype"] == "f64"
    assert d["vs_baseline"] is None and "workload" in d["config"] and "l2" in d["config"]
    assert d["warmup"] >= 3 and d["steps"] >= 1 and d["n_gpus"] == 1
    # value is consistent with the timing it was derived from
    assert abs(d["value"] - d["config"]["objects_per_gpu"] / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert 0 < r["frac"] < 1 and r["traffic"] is not None
    e = d["e2e"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(e) and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert e["value"] != d["value"]                       # measured separately, not a copy of the device-timed value
    c = d["cpu_baseline"]
    assert {"value", "unit", "cores", "kind", "sample"} <= set(c) and c["kind"] in ("port", "reference") and c["cores"] >= 1
    assert d["gpu_launches"] >= 5 * d["steps"]
    k = d["clocks"]
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(k)
    assert not any("slowdown" in str(x) for x in k["reasons"])


def test_other_committed_lines_are_consistent():
    for name in ("r01_bench_c4_1M_1gpu.json", "r01_bench_c3_device_rng.json"):
        d = _line(name)
        assert BASE_KEYS <= set(d), (name, BASE_KEYS - set(d))
        assert d["roofline"]["frac"] > 0 and d["clocks"]["sm_mhz"] > 0
    for w in ("c2", "c4", "c3"):
        vals = [_line(f"r01_scale_{w}_n{n}.json")["value"] for n in (1, 2, 4, 8)]
        assert vals == sorted(vals) and vals[3] > 5 * vals[0], (w, vals)
    ref = _line("r01_bench_reference_arm.json")
    assert ref["impl"] == "reference" and ref["e2e"]["h2d_bytes_per_step"] == 0 and ref["cpu_baseline"]["value"] == ref["value"]


def test_reference_arm_runs_on_the_host():
    out = subprocess.run([sys.executable, os.path.join(H.ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--objects", "2000"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-500:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] == "port"
