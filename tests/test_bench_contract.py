"""The bench line's contract (the driver parses it): checked on the committed captures under profiles/ -- the lines bench.py printed on
the B200 boxes -- and on bench.py's pure helpers.  No GPU."""
import glob
import importlib.util
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LINES = sorted(glob.glob(os.path.join(ROOT, "profiles", "r1j_bench*.json")) + glob.glob(os.path.join(ROOT, "profiles", "r1i_bench*.json")))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


@pytest.mark.parametrize("path", LINES, ids=[os.path.basename(p) for p in LINES])
def test_committed_line_keeps_the_contract(path):
    d = json.load(open(path))
    for k, t in (("metric", str), ("value", float), ("unit", str), ("n_gpus", int), ("steps", int), ("warmup", int), ("ms_per_step", float),
                 ("higher_is_better", bool), ("scaling", str), ("dtype", str), ("data", str), ("config", dict)):
        assert isinstance(d[k], t), k
    assert d["unit"] == "Mrays/s" and d["higher_is_better"] is True and d["warmup"] >= 1 and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    assert abs(d["value"] - d["config"].get("rays_per_step", d["value"] * d["ms_per_step"] * 1e3) / (d["ms_per_step"] * 1e-3) / 1e6) < 1e-6 * d["value"] or d.get("impl") == "reference"
    e = d["e2e"]
    if e is not None:
        assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(e)
    if d.get("impl") == "reference":
        assert d["cpu_baseline"]["kind"] in ("port", "reference") and e["h2d_bytes_per_step"] == 0 and e["value"] == d["value"]
        return
    assert d["warmup"] >= 3 and d["gpu_launches"] > 0
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and "traffic" in r
    assert d["clocks"]["reasons"] == [] or set(d["clocks"]["reasons"]) <= {"sw_power_cap"}
    if d["n_gpus"] == 1 and d["cpu_baseline"] is not None:
        c = d["cpu_baseline"]
        assert {"value", "unit", "cores", "kind", "sample"} <= set(c) and c["kind"] == "port" and c["cores"] >= 1
        assert e is not None and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] < d["value"]
        p = d.get("parity")
        if p is not None:                                   # captures since r1i carry the full-size parity block
            assert p["bit_exact"] is True and all(p["planes_bit_exact"].values())


def test_there_are_captures_for_every_config_and_gpu_count():
    names = {os.path.basename(p) for p in LINES}
    for want in ("r1j_bench.json", "r1i_bench_cfg1.json", "r1i_bench_cfg2.json", "r1i_bench_cfg4.json", "r1i_bench_cfg5.json",
                 "r1i_bench_cfg3_n2.json", "r1i_bench_cfg3_n4.json", "r1i_bench_cfg3_n8.json", "r1j_bench_cfg3_n8.json", "r1i_bench_reference_arm.json"):
        assert want in names, want


def test_issue_roofline_helper():
    m = _bench()
    r = m.issue_roofline("k_ambient<1>", 4.77, 1965.0, 148)
    assert r is not None and 0.5 < r["frac"] < 1.0 and abs(r["peak_warp_inst_per_s"] - 148 * 4 * 1965e6) < 1.0
    assert m.issue_roofline("k_ambient", 4.77, None, 148) is None and m.issue_roofline("no_such_kernel", 1.0, 1965.0, 148) is None


def test_algorithmic_bytes_follow_survey_8d():
    m = _bench()
    st = dict(rays=17, steps=1000, pixels=1)
    # ambient: probes * 1 B + lit pixels * (depth 4 + normal 4 + noise 4 * n_ao ... as DESIGN 5.1 states: 16 + 4 * n_ao)
    assert m.algorithmic_bytes("ambient", st, 16) == 1000 + 1 * (16 + 4 * 16)
