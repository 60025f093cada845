"""Results file of a model run from the device recorder: the variables of the reference's `results.nc`
(core/src/base/ResultWriter.cpp:8-87, written by ModelHydro::DumpOutputs -> Logger::DumpOutputs) under the same names, with
the same dimension order and attributes, and a reader with the surface of python/src/hydrobricks/results.py (`Results`).

Two containers are written side by side: `results.nc` — netCDF classic (64-bit offset) through scipy.io, the one netCDF
writer in this image; xarray, ncdump or the netCDF4 package open it like the reference's file — and `results.npz`, the
same arrays for numpy alone.

      time                       [time]                       float64, days since 1858-11-17 (MJD)
      hydro_units_ids            [hydro_units]                int32
      hydro_units_structure_ids  [hydro_units]                int32
      hydro_units_areas          [hydro_units]                float64
      sub_basin_values           [aggregated_values][time]    float64 (mm)
      hydro_units_values         [distributed_values][hydro_units][time]
      land_cover_fractions       [land_covers][hydro_units][time]        (only when fractions were recorded)
      labels_aggregated, labels_distributed, labels_land_covers          global attributes

Stated differences: (1) the engine's sub-basin recorder holds the outlet discharge only, so `labels_aggregated` is
["outlet"] (the reference also logs the contents and outflows of the two glacier-area stores when log_all is set);
(2) the classic format has no string-array attributes (the reference's netCDF-4 file stores the labels as NC_STRING
arrays, ResultWriter.cpp:76-79): the label lists are comma-joined text attributes (no label contains a comma), and the
variables are not compressed.
"""
from __future__ import annotations

from pathlib import Path

import numpy as np

FILE_NAME = "results.npz"
NC_FILE_NAME = "results.nc"
_LONG_NAMES = {  # ResultWriter.cpp:41-72: variable -> (dimensions, long_name, units)
    "time": (("time",), "time", "days since 1858-11-17 00:00:00.0"),
    "hydro_units_ids": (("hydro_units",), "hydrological units ids", None),
    "hydro_units_structure_ids": (("hydro_units",), "model structure id used by each hydrological unit", None),
    "hydro_units_areas": (("hydro_units",), "hydrological units areas", None),
    "sub_basin_values": (("aggregated_values", "time"), "aggregated values over the sub basin", "mm"),
    "hydro_units_values": (("distributed_values", "hydro_units", "time"), "values for each hydrological units", "mm"),
    "land_cover_fractions": (("land_covers", "hydro_units", "time"), "land cover fractions for each hydrological units",
                             "percent"),
}


def write_results(model, path, structure_ids=None, netcdf=True):
    """ModelHydro::DumpOutputs(path) for either host module's ModelHydro (after run()): writes <path>/results.nc (unless
    netcdf=False) and <path>/results.npz and returns True; False when the directory does not exist
    (ResultWriter.cpp:13-16)."""
    directory = Path(path)
    if not directory.is_dir():
        return False
    outlet = np.asarray(model.get_outlet_discharge(), dtype=np.float64)
    n_steps = len(outlet)
    ids = np.asarray(model.get_hydro_unit_ids(), dtype=np.int32)
    areas = np.asarray(model.get_hydro_unit_areas(), dtype=np.float64)
    labels = list(model.get_recorded_hydro_unit_labels())
    values = np.empty((len(labels), len(ids), n_steps), dtype=np.float64)
    for i, label in enumerate(labels):
        values[i] = np.asarray(model.get_hydro_unit_values(label)).T  # [T x U] -> [U x T]
    arrays = {
        "time": _time_axis(model, n_steps),
        "hydro_units_ids": ids,
        "hydro_units_structure_ids": np.asarray(structure_ids if structure_ids is not None
                                                else _structure_ids(model, len(ids)), dtype=np.int32),
        "hydro_units_areas": areas,
        "sub_basin_values": outlet[None, :],
        "hydro_units_values": values,
        "labels_aggregated": np.array(["outlet"]),
        "labels_distributed": np.array(labels, dtype=str),
    }
    fraction_labels = list(model.get_recorded_hydro_unit_fraction_labels())
    if fraction_labels:
        fractions = np.empty((len(fraction_labels), len(ids), n_steps), dtype=np.float64)
        for i, label in enumerate(fraction_labels):
            fractions[i] = np.asarray(model.get_hydro_unit_fractions(label)).T
        arrays["land_cover_fractions"] = fractions
        arrays["labels_land_covers"] = np.array(fraction_labels, dtype=str)
    np.savez_compressed(directory / FILE_NAME, **arrays)
    if netcdf:
        _write_netcdf(directory / NC_FILE_NAME, arrays)
    return True


def _write_netcdf(path, arrays):
    """ResultWriter.cpp:21-87 in netCDF classic, 64-bit offset (a variable may hold up to 4 GiB)."""
    from scipy.io import netcdf_file

    with netcdf_file(str(path), "w", version=2) as nc:
        for name, (dims, _, _) in _LONG_NAMES.items():
            if name in arrays:
                for dim, size in zip(dims, arrays[name].shape):
                    if dim not in nc.dimensions:
                        nc.createDimension(dim, int(size))
        for name, (dims, long_name, units) in _LONG_NAMES.items():
            if name not in arrays:
                continue
            var = nc.createVariable(name, "i" if arrays[name].dtype.kind == "i" else "d", dims)
            var[:] = arrays[name]
            var.long_name = long_name
            if units:
                var.units = units
        for name in ("labels_aggregated", "labels_distributed", "labels_land_covers"):
            if name in arrays:
                labels = [str(x) for x in arrays[name]]
                if any("," in x for x in labels):
                    raise ValueError(f"a label with a comma cannot be stored in {NC_FILE_NAME}: {labels}")
                setattr(nc, name, ",".join(labels))


def _read_netcdf(path):
    from scipy.io import netcdf_file

    with netcdf_file(str(path), "r", mmap=False) as nc:
        out = {name: np.array(var[:]) for name, var in nc.variables.items()}
        for name in ("labels_aggregated", "labels_distributed", "labels_land_covers"):
            text = getattr(nc, name, None)
            if text is not None:
                text = text.decode() if isinstance(text, bytes) else str(text)
                out[name] = np.array(text.split(",") if text else [], dtype=str)
    return out


def _time_axis(model, n_steps):
    start = getattr(model, "start_mjd", None)
    start = start() if callable(start) else start
    if start is None:
        start = getattr(model, "_mjd0", 0.0)
    step = getattr(model, "time_step_days", None)
    step = step() if callable(step) else step
    if step is None:
        cs = getattr(model, "_cs", None)
        step = cs.hbk.time_step_days if cs is not None else 1.0
    return float(start) + float(step) * np.arange(n_steps, dtype=np.float64)


def _structure_ids(model, n_units):
    get = getattr(model, "get_hydro_unit_structure_ids", None)
    return get() if get is not None else np.ones(n_units, dtype=np.int32)


class Results:
    """python/src/hydrobricks/results.py `Results` on a results.nc / results.npz: the same attribute and method names,
    numpy instead of xarray. Dates are 'YYYY-MM-DD' strings; time selections return views [hydro_units x time] (or [hydro_units] for a
    single date), like the reference's `.sel(time=slice(...))`."""

    def __init__(self, filename):
        p = Path(filename)
        if p.is_dir():
            p = p / FILE_NAME if (p / FILE_NAME).exists() else p / NC_FILE_NAME
        if p.suffix == ".nc":
            self.results = _read_netcdf(p)
        else:
            with np.load(p) as z:
                self.results = {k: z[k] for k in z.files}
        self.labels_distributed = [str(x) for x in self.results["labels_distributed"]]
        self.labels_aggregated = [str(x) for x in self.results["labels_aggregated"]]
        self.labels_land_cover = [str(x) for x in self.results["labels_land_covers"]] \
            if "labels_land_covers" in self.results else None
        self.hydro_units_ids = self.results["hydro_units_ids"]
        self._dates = np.datetime64("1858-11-17") + np.floor(self.results["time"]).astype("timedelta64[D]")

    def close(self):
        self.results = None

    def __enter__(self):
        return self

    def __exit__(self, exc_type, exc_val, exc_tb):
        self.close()
        return False

    def list_hydro_units_components(self):
        print("Hydro units components:")
        for label in self.labels_distributed:
            print("- " + label)

    def list_sub_basin_components(self):
        print("Sub basins components:")
        for label in self.labels_aggregated:
            print("- " + label)

    def get_hydro_units_structure_ids(self):
        return self.results["hydro_units_structure_ids"]

    def get_time_array(self, start_date, end_date):
        return self._dates[self._slice(start_date, end_date)]

    def get_sub_basin_values(self, component, start_date=None, end_date=None):
        if component not in self.labels_aggregated:
            raise ValueError(f"The component '{component}' was not found in the results.")
        return self._select_time(self.results["sub_basin_values"][self.labels_aggregated.index(component)],
                                 start_date, end_date)

    def get_hydro_units_values(self, component, start_date=None, end_date=None):
        i_component, _ = self._resolve_component_label(component)
        return self._select_time(self.results["hydro_units_values"][i_component], start_date, end_date)

    def get_land_cover_areas(self, land_cover, start_date=None, end_date=None):
        if not self.labels_land_cover or land_cover not in self.labels_land_cover:
            raise ValueError(f"The land cover '{land_cover}' was not found in the results.")
        fractions = self.results["land_cover_fractions"][self.labels_land_cover.index(land_cover)]
        areas = fractions * self.results["hydro_units_areas"][:, None]
        return self._select_time(areas, start_date, end_date)

    def get_mean_hydro_units_values(self, land_cover, component, start_date=None, end_date=None):
        values = self.get_hydro_units_values(component, start_date, end_date)
        lc_areas = self.get_land_cover_areas(land_cover, start_date, end_date)
        return self._area_weighted_nanmean(values, lc_areas, axis=0)

    def get_mean_swe(self, start_date=None, end_date=None):
        lc_swe, lc_areas = self._stack_land_cover_swe(start_date, end_date)
        lc_swe = lc_swe.reshape(-1, *lc_swe.shape[2:])
        lc_areas = lc_areas.reshape(-1, *lc_areas.shape[2:])
        return self._area_weighted_nanmean(lc_swe, lc_areas, axis=0)

    def get_total_swe(self, start_date=None, end_date=None):
        lc_swe, lc_areas = self._stack_land_cover_swe(start_date, end_date)
        return self._area_weighted_nanmean(lc_swe, lc_areas, axis=0)

    # -- helpers ------------------------------------------------------------------------------------------------
    def _stack_land_cover_swe(self, start_date, end_date):
        if not self.labels_land_cover:
            raise ValueError("The land cover fractions were not recorded.")
        swe, areas = [], []
        for cover in self.labels_land_cover:
            label = f"{cover}_snowpack:snow_content"
            area = self.get_land_cover_areas(cover, start_date, end_date)
            if label in self.labels_distributed:
                swe.append(np.nan_to_num(self.get_hydro_units_values(label, start_date, end_date)))
            else:  # a cover without a snowpack contributes zero SWE over its area
                swe.append(np.zeros_like(area))
            areas.append(area)
        return np.stack(swe), np.stack(areas)

    def _slice(self, start_date, end_date):
        if start_date is None and end_date is None:
            return slice(None)
        i0 = 0 if start_date is None else int(np.searchsorted(self._dates, np.datetime64(start_date)))
        if i0 >= len(self._dates) or (start_date is not None and self._dates[i0] != np.datetime64(start_date)):
            raise KeyError(f"date {start_date} is not in the time series")
        if end_date is None:  # a single-date snapshot (results.py: `.sel(time=start_date)`)
            return i0
        i1 = int(np.searchsorted(self._dates, np.datetime64(end_date), side="right"))
        return slice(i0, i1)

    def _select_time(self, data, start_date, end_date):
        return data[..., self._slice(start_date, end_date)]

    @staticmethod
    def _area_weighted_nanmean(values, areas, axis):
        weights = np.where(np.isnan(values), 0.0, areas)
        total = weights.sum(axis=axis)
        with np.errstate(invalid="ignore", divide="ignore"):
            return np.where(total > 0, np.nansum(values * weights, axis=axis) / total, np.nan)

    def _has_distributed_component(self, component):
        return component in self.labels_distributed

    def _resolve_component_label(self, component):
        """results.py:431-455: the exact label, or a unique suffix match ("slow_reservoir:water_content" for
        "ground_slow_reservoir:water_content")."""
        labels = self.labels_distributed
        if component in labels:
            return labels.index(component), component
        matches = [label for label in labels if label.endswith(component)]
        if len(matches) == 1:
            return labels.index(matches[0]), matches[0]
        if len(matches) > 1:
            raise ValueError(f"Component '{component}' is ambiguous. Matching labels: {matches}.")
        raise ValueError(f"Component '{component}' not found in distributed results. "
                         f"Available components: {', '.join(labels)}.")
