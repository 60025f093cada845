"""Load the UNMODIFIED reference Python package (/root/reference/python/src/hydrobricks) on top of a
chosen `_hydrobricks` extension module. TEST/GOLDEN-GENERATION INFRASTRUCTURE, build container only
(/root/reference does not exist on the GPU box).

    hb = load_reference_python(backend="ref")      # oracle/_ref/_hydrobricks*.so = real reference core
    hb = load_reference_python(backend="b200")     # hydrobricks_b200/_hydrobricks*.so = this repo's engine

HydroErr and cftime are not installed in this image; tiny stand-ins are injected (HydroErr: nse and
kge_2012 from their published formulas; cftime.num2date for the 'days since' axis the forcing loader
uses). Both are outside the ModelHydro::Run() path.
"""
import importlib
import importlib.util
import sys
import types
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
REF_PY = Path("/root/reference/python/src")


def _stub_hydroerr():
    m = types.ModuleType("HydroErr")

    def _clean(sim, obs):
        sim = np.asarray(sim, dtype=float)
        obs = np.asarray(obs, dtype=float)
        ok = ~(np.isnan(sim) | np.isnan(obs))
        return sim[ok], obs[ok]

    def nse(simulated_array, observed_array, **_):
        s, o = _clean(simulated_array, observed_array)
        return 1 - np.sum((s - o) ** 2) / np.sum((o - o.mean()) ** 2)

    def kge_2012(simulated_array, observed_array, s=(1, 1, 1), **_):
        sim, obs = _clean(simulated_array, observed_array)
        r = np.corrcoef(sim, obs)[0, 1]
        beta = sim.mean() / obs.mean()
        gamma = (sim.std(ddof=1) / sim.mean()) / (obs.std(ddof=1) / obs.mean())
        return 1 - np.sqrt((s[0] * (r - 1)) ** 2 + (s[1] * (gamma - 1)) ** 2 + (s[2] * (beta - 1)) ** 2)

    m.nse = nse
    m.kge_2012 = kge_2012
    m.__all__ = ["nse", "kge_2012"]
    return m


def _stub_cftime():
    m = types.ModuleType("cftime")

    def num2date(times, units, calendar="standard", **_):
        import pandas as pd

        unit, _, origin = units.partition(" since ")
        return pd.to_datetime(np.asarray(times), unit={"days": "D", "hours": "h"}.get(unit.strip(), "D"),
                              origin=pd.Timestamp(origin.strip())).to_pydatetime()

    m.num2date = num2date
    return m


class _Anything(types.ModuleType):
    """Placeholder for plotting/GIS modules the reference imports at package import time."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        sub = _Anything(f"{self.__name__}.{name}")
        setattr(self, name, sub)
        return sub

    def __call__(self, *a, **k):
        return self


def _stub_missing(names):
    for name in names:
        try:
            importlib.import_module(name)
        except ImportError:
            parts = name.split(".")
            for i in range(1, len(parts) + 1):
                mod = _Anything(".".join(parts[:i]))
                mod.__path__ = []  # behave like a package so submodule imports resolve
                sys.modules.setdefault(".".join(parts[:i]), mod)


def load_extension(backend):
    if backend == "ref":
        cands = sorted((REPO / "oracle" / "_ref").glob("_hydrobricks*.so"))
    elif backend == "b200":
        if str(REPO) not in sys.path:
            sys.path.insert(0, str(REPO))
        return importlib.import_module("hydrobricks_b200._hydrobricks")
    elif backend == "b200-native":
        if str(REPO) not in sys.path:
            sys.path.insert(0, str(REPO))
        return importlib.import_module("hydrobricks_b200.native._hydrobricks")
    else:
        raise ValueError(backend)
    if not cands:
        raise ImportError(f"no _hydrobricks extension built for backend {backend!r}")
    spec = importlib.util.spec_from_file_location("_hydrobricks", cands[0])
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load_reference_python(backend="ref"):
    if not REF_PY.exists():
        raise ImportError("/root/reference is not available here")
    for name in [n for n in sys.modules if n == "hydrobricks" or n.startswith("hydrobricks.")]:
        del sys.modules[name]
    sys.modules.setdefault("HydroErr", _stub_hydroerr())
    sys.modules.setdefault("cftime", _stub_cftime())
    _stub_missing(["matplotlib", "matplotlib.pyplot", "matplotlib.dates", "matplotlib.colors",
                   "matplotlib.patches", "matplotlib.lines", "matplotlib.ticker", "matplotlib.gridspec",
                   "matplotlib.cm", "matplotlib.figure", "matplotlib.axes", "matplotlib.animation"])
    ext = load_extension(backend)
    if str(REF_PY) not in sys.path:
        sys.path.insert(0, str(REF_PY))
    sys.modules["hydrobricks._hydrobricks"] = ext
    hb = importlib.import_module("hydrobricks")
    return hb
