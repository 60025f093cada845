"""results.npz writer / reader (hydrobricks_b200/results.py: the variables of the reference's results.nc,
core/src/base/ResultWriter.cpp:8-87, and the surface of python/src/hydrobricks/results.py) on a stand-in model object."""
import numpy as np
import pytest

from hydrobricks_b200 import results


class FakeModel:
    """The getters write_results uses (ModelHydro surface, Binding.cpp:308-360)."""

    def __init__(self, T=20, U=3):
        rng = np.random.default_rng(1)
        self.T, self.U = T, U
        self.q = rng.uniform(0, 5, T)
        self.labels = ["ground_snowpack:snow_content", "glacier_snowpack:snow_content", "slow_reservoir:water_content"]
        self.values = {label: rng.uniform(0, 100, (T, U)) for label in self.labels}
        self.values["glacier_snowpack:snow_content"][:, 2] = np.nan  # a unit without the glacier bricks
        self.fractions = {"ground": np.tile([0.4, 0.7, 1.0], (T, 1)), "glacier": np.tile([0.6, 0.3, 0.0], (T, 1))}

    def get_outlet_discharge(self):
        return self.q

    def get_hydro_unit_ids(self):
        return [7, 8, 9]

    def get_hydro_unit_areas(self):
        return [1.0e6, 2.0e6, 3.0e6]

    def get_recorded_hydro_unit_labels(self):
        return self.labels

    def get_hydro_unit_values(self, label):
        return self.values[label]

    def get_recorded_hydro_unit_fraction_labels(self):
        return list(self.fractions)

    def get_hydro_unit_fractions(self, label):
        return self.fractions[label]

    def get_hydro_unit_structure_ids(self):
        return [2, 2, 1]

    def start_mjd(self):
        return 44605.0  # 1981-01-01

    def time_step_days(self):
        return 1.0


def test_write_then_read(tmp_path):
    m = FakeModel()
    assert results.write_results(m, tmp_path)
    assert not results.write_results(m, tmp_path / "missing")  # ResultWriter.cpp:13-16
    with results.Results(tmp_path / "results.npz") as r:
        assert r.labels_aggregated == ["outlet"] and r.labels_distributed == m.labels
        assert r.labels_land_cover == ["ground", "glacier"]
        assert list(r.hydro_units_ids) == [7, 8, 9] and list(r.get_hydro_units_structure_ids()) == [2, 2, 1]
        # the reference's dimension order: [distributed_values][hydro_units][time]
        assert r.results["hydro_units_values"].shape == (3, 3, 20) and r.results["sub_basin_values"].shape == (1, 20)
        assert np.array_equal(r.get_sub_basin_values("outlet"), m.q)
        v = r.get_hydro_units_values("slow_reservoir:water_content")
        assert np.array_equal(v, m.values["slow_reservoir:water_content"].T)
        # unique suffix match (results.py:431-455), ambiguity refused
        assert np.array_equal(r.get_hydro_units_values("reservoir:water_content"), v)
        with pytest.raises(ValueError, match="ambiguous"):
            r.get_hydro_units_values("snow_content")
        with pytest.raises(ValueError, match="not found"):
            r.get_hydro_units_values("canopy")
        # time selection: slice and single-date snapshot
        sel = r.get_hydro_units_values("slow_reservoir:water_content", "1981-01-03", "1981-01-05")
        assert np.array_equal(sel, v[:, 2:5])
        assert np.array_equal(r.get_hydro_units_values("slow_reservoir:water_content", "1981-01-03"), v[:, 2])
        assert list(r.get_time_array("1981-01-01", "1981-01-02").astype(str)) == ["1981-01-01", "1981-01-02"]
        with pytest.raises(KeyError):
            r.get_hydro_units_values("slow_reservoir:water_content", "1979-01-01", "1981-01-05")
        # land-cover areas and area-weighted means
        areas = r.get_land_cover_areas("glacier")
        assert np.allclose(areas[:, 0], [0.6e6, 0.6e6, 0.0])
        mean = r.get_mean_hydro_units_values("glacier", "glacier_snowpack:snow_content")
        s = m.values["glacier_snowpack:snow_content"]
        assert np.allclose(mean, (s[:, 0] * 0.6e6 + s[:, 1] * 0.6e6) / 1.2e6)
        total = r.get_total_swe()  # per unit: the land covers' snowpacks weighted by their areas
        g = m.values["ground_snowpack:snow_content"]
        assert np.allclose(total[0], g[:, 0] * 0.4 + s[:, 0] * 0.6)
        assert np.allclose(total[2], g[:, 2])
        mean_swe = r.get_mean_swe()
        w = np.array([0.4e6, 1.4e6, 3.0e6, 0.6e6, 0.6e6, 0.0])
        stacked = np.concatenate([g.T, np.nan_to_num(s.T)])
        assert np.allclose(mean_swe, (stacked * w[:, None]).sum(0) / w.sum())


def test_results_nc_holds_the_reference_layout(tmp_path):
    """results.nc (netCDF classic through scipy.io): dimensions, variables, attributes and units of ResultWriter.cpp:21-87,
    the same numbers as results.npz, and the Results reader gives the same answers on either file."""
    from scipy.io import netcdf_file

    m = FakeModel()
    assert results.write_results(m, tmp_path)
    with netcdf_file(str(tmp_path / "results.nc"), "r", mmap=False) as nc:
        assert {k: v for k, v in nc.dimensions.items()} == {"time": 20, "hydro_units": 3, "aggregated_values": 1,
                                                            "distributed_values": 3, "land_covers": 2}
        assert nc.variables["time"].units == b"days since 1858-11-17 00:00:00.0"
        assert nc.variables["time"].long_name == b"time" and nc.variables["time"][0] == 44605.0
        assert nc.variables["hydro_units_values"].dimensions == ("distributed_values", "hydro_units", "time")
        assert nc.variables["hydro_units_values"].units == b"mm" and nc.variables["sub_basin_values"].units == b"mm"
        assert nc.variables["land_cover_fractions"].dimensions == ("land_covers", "hydro_units", "time")
        assert nc.variables["land_cover_fractions"].units == b"percent"
        assert nc.variables["hydro_units_ids"].typecode() == "i" and list(nc.variables["hydro_units_ids"][:]) == [7, 8, 9]
        assert nc.variables["hydro_units_structure_ids"].long_name == b"model structure id used by each hydrological unit"
        assert nc.labels_distributed.decode().split(",") == m.labels and nc.labels_aggregated == b"outlet"
    with results.Results(tmp_path / "results.nc") as a, results.Results(tmp_path / "results.npz") as b:
        assert a.labels_distributed == b.labels_distributed and a.labels_land_cover == b.labels_land_cover
        for key in b.results:
            if not key.startswith("labels_"):
                assert np.array_equal(a.results[key], b.results[key], equal_nan=True), key
        assert np.array_equal(a.get_sub_basin_values("outlet"), m.q)
        assert np.array_equal(a.get_mean_swe(), b.get_mean_swe())
        assert np.array_equal(a.get_hydro_units_values("reservoir:water_content", "1981-01-03"),
                              b.get_hydro_units_values("reservoir:water_content", "1981-01-03"))
    # a directory is enough; the .npz can be left out
    (tmp_path / "only_nc").mkdir()
    assert results.write_results(m, tmp_path / "only_nc")
    (tmp_path / "only_nc" / "results.npz").unlink()
    with results.Results(tmp_path / "only_nc") as r:
        assert r.labels_aggregated == ["outlet"]
    (tmp_path / "no_nc").mkdir()
    assert results.write_results(m, tmp_path / "no_nc", netcdf=False) and not (tmp_path / "no_nc" / "results.nc").exists()
