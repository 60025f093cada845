"""bench.py on the CPU: the synthetic workload generators and the reference arm's JSON contract
(`--impl reference` runs the reference's own CPU implementation; here with a tiny sample so that it takes seconds)."""
import json
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import bench  # noqa: E402


def test_synthetic_inputs_are_deterministic_and_in_range():
    p1, t1, e1 = bench.station_series(400)
    p2, t2, e2 = bench.station_series(400)
    assert np.array_equal(p1, p2) and np.array_equal(t1, t2) and np.array_equal(e1, e2)
    assert (p1 >= 0).all() and (e1 >= 0).all() and 0.3 < (p1 > 0).mean() < 0.5  # rain on ~40 % of the days
    ps = bench.parameter_sets(5000)
    assert (ps["prec_t_start"] < ps["prec_t_end"]).all()
    assert 2 <= ps["a_snow"].min() and ps["a_snow"].max() <= 20 and ps["A"].max() <= 3000 and ps["beta"].min() >= 100
    units = bench.hydro_units(50, glacier=True)
    assert len(units) == 50 and all(abs(sum(f for _, f in u["land_covers"]) - 1.0) < 1e-12 for u in units)
    assert max(dict(u["land_covers"])["glacier"] for u in units) <= 0.8
    assert {"c2", "c3", "c4", "c5"} <= set(bench.WORKLOADS)
    assert bench.WORKLOADS["c2"]["sets"] == 100000 and bench.WORKLOADS["c2"]["timesteps"] == 10957


def test_c4_tables_and_changes():
    units = sorted(bench.hydro_units(50, glacier=True), key=lambda u: -u["elevation"])
    gl, areas, volumes = bench.c4_tables(units)
    a0 = np.array([units[i]["area"] * dict(units[i]["land_covers"])["glacier"] for i in gl])
    assert areas.shape == volumes.shape == (11, len(gl)) and np.array_equal(areas[0], a0)  # Init compares bit for bit
    assert (areas[-1] == 0).all() and (np.diff(areas, axis=0) <= 1e-9).all() and (volumes >= 0).all()
    changes = bench.c4_changes(units, gl, areas, 18262)
    assert len(changes) == 49 and all(0 < area <= units_area for (_, uid, area), units_area in
                                      zip(changes, [next(u["area"] for u in units if u["id"] == c[1]) for c in changes]))
    dates = [c[0] for c in changes]
    assert dates == sorted(dates) and dates[0] > bench.mjd_axis(1)[0]


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--units", "6", "--timesteps", "120", "--ref-sets-per-core", "1"],
                         capture_output=True, text=True, timeout=600)
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, out.stderr[-500:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "hru_timesteps_per_s" and d["unit"] == "HRU*timesteps/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1 and d["warmup"] == 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] and "sample" in d["cpu_baseline"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("C2") and d["dtype"] == "f64" and d["vs_baseline"] is None
