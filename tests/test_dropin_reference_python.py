"""Drop-in check of the host mirror under the UNMODIFIED reference Python package (only where /root/reference
exists, i.e. the build container; no GPU there, so this covers everything up to engine creation): the reference's
hb.models.Socont must build, through hydrobricks_b200._hydrobricks, exactly the structure it builds through the
compiled reference module, and setup() must reach the engine (and fail loudly for lack of a CUDA device)."""
import copy
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "tools"))

pytestmark = pytest.mark.skipif(not Path("/root/reference/python/src/hydrobricks").exists() or
                                not list((ROOT / "oracle" / "_ref").glob("_hydrobricks*.so")),
                                reason="reference sources / compiled reference not available")


def _structure(backend, **kw):
    import refpy

    refpy.load_reference_python(backend)
    import hydrobricks.models as models

    socont = models.Socont(**kw)
    st = socont.settings.settings.get_structure()
    st = copy.deepcopy(st)
    for s in st:
        s.pop("log_items", None)
    return st, socont


@pytest.mark.parametrize("kw", [
    dict(soil_storage_nb=1, record_all=True),  # the reference layer's default solver: crank_nicolson (model.py:80)
    dict(solver="rk4", soil_storage_nb=1, surface_runoff="socont_runoff", record_all=True),
    dict(solver="heun_explicit", soil_storage_nb=2, surface_runoff="linear_storage", record_all=False),
    dict(solver="rk4", soil_storage_nb=2, land_cover_names=["ground", "glacier"], land_cover_types=["ground", "glacier"],
         record_all=True),
    dict(solver="euler_explicit", soil_storage_nb=1, land_cover_names=["ground", "glacier"],
         land_cover_types=["ground", "glacier"], glacier_infinite_storage=False, record_all=True),
    dict(solver="rk4", soil_storage_nb=1, land_cover_names=["ground", "glacier"], land_cover_types=["ground", "glacier"],
         glacier_infinite_storage=False, record_all=True, snow_ice_transformation="transform:snow_ice_swat"),
    dict(solver="rk4", soil_storage_nb=2, land_cover_names=["ground", "glacier"], land_cover_types=["ground", "glacier"],
         record_all=False, snow_ice_transformation="transform:snow_ice_constant"),
    dict(solver="rk4", soil_storage_nb=1, record_all=True, snow_sublimation_process="sublimation:constant"),
    dict(solver="heun_explicit", soil_storage_nb=2, land_cover_names=["ground", "glacier"],
         land_cover_types=["ground", "glacier"], record_all=True, snow_ice_transformation="transform:snow_ice_swat",
         snow_sublimation_process="sublimation:pet"),
    # round 2: the other melt processes, the lateral snow slide (with and without the glacier snowpacks), two glacier covers
    dict(solver="rk4", soil_storage_nb=1, land_cover_names=["ground", "glacier"], land_cover_types=["ground", "glacier"],
         record_all=True, snow_melt_process="melt:degree_day_aspect"),
    dict(solver="heun_explicit", soil_storage_nb=2, land_cover_names=["ground", "glacier"],
         land_cover_types=["ground", "glacier"], record_all=True, snow_melt_process="melt:temperature_index"),
    dict(solver="rk4", soil_storage_nb=1, land_cover_names=["ground", "glacier"], land_cover_types=["ground", "glacier"],
         record_all=True, snow_redistribution="transport:snow_slide"),
    dict(solver="rk4", soil_storage_nb=2, land_cover_names=["ground", "glacier"], land_cover_types=["ground", "glacier"],
         record_all=True, snow_redistribution="transport:snow_slide", snow_ice_transformation="transform:snow_ice_swat",
         glacier_infinite_storage=False),
    dict(solver="rk4", soil_storage_nb=1, land_cover_names=["ground", "glacier_ice", "glacier_debris"],
         land_cover_types=["ground", "glacier", "glacier"], glacier_infinite_storage=False, record_all=True),
])
@pytest.mark.parametrize("backend", ["b200", "b200-native"])
def test_reference_socont_builds_same_structure_on_both_backends(kw, backend):
    ref_st, _ = _structure("ref", **kw)
    b200_st, socont = _structure(backend, **kw)
    assert b200_st == ref_st
    # and the parameter plumbing of the reference layer works on the mirror
    params = socont.generate_parameters()
    params.set_values({"A": 200})
    for _, p in params.get_model_parameters().iterrows():
        if p["value"] is None:
            continue
        assert socont.settings.set_parameter_value(p["component"], p["name"], p["value"])


@pytest.mark.parametrize("backend", ["b200", "b200-native"])
def test_reference_setup_reaches_the_engine(backend):
    import refpy
    import torch

    hb = refpy.load_reference_python(backend)
    import hydrobricks.models as models

    socont = models.Socont(solver="rk4", soil_storage_nb=1, surface_runoff="linear_storage")
    hu = hb.HydroUnits()
    hu.load_from_csv("/root/reference/tests/files/catchments/ch_sitter_appenzell/hydro_units_elevation.csv",
                     column_elevation="elevation", column_area="area")
    if torch.cuda.is_available():
        socont.setup(spatial_structure=hu, output_path="/tmp", start_date="1981-01-01", end_date="1981-12-31")
    else:
        with pytest.raises(Exception) as err:
            socont.setup(spatial_structure=hu, output_path="/tmp", start_date="1981-01-01", end_date="1981-12-31")
        assert "CUDA" in str(err.value) or "hbk" in str(err.value)


def test_the_reference_own_python_tests_pass_on_the_host_module():
    """The reference's OWN test files that stop before a model run (parameters, binding surface, hydro units, ET process
    structures, periods), with its
    extension module replaced by this repo's: every test passes (the two skipped ones want SPOTPY)."""
    import re
    import subprocess

    out = subprocess.run([sys.executable, str(ROOT / "tools" / "run_reference_tests.py"), "b200", "test_parameters.py",
                          "test_binding.py", "test_hydro_units.py", "test_et_processes.py", "test_periods.py"], capture_output=True, text=True, timeout=600)
    tail = out.stdout.strip().splitlines()[-1]
    assert out.returncode == 0 and "failed" not in tail and "error" not in tail, out.stdout[-2000:]
    assert int(re.search(r"(\d+) passed", tail).group(1)) >= 81
