"""Helpers shared by the parity tests: load tests/golden/*.npz bundles (made by tools/make_goldens.py from
the unmodified reference) and run the oracle restatement on them."""
import json
from pathlib import Path

import numpy as np

GOLDEN = Path(__file__).resolve().parent / "golden"


def bundles(prefix=""):
    return sorted(p.stem for p in GOLDEN.glob(f"{prefix}*.npz"))


def load(name):
    z = np.load(GOLDEN / f"{name}.npz")
    meta = json.loads(bytes(z["meta"]).decode())
    forcing = {k[len("forcing_"):]: z[k] for k in z.files if k.startswith("forcing_")}
    records = {label: z[f"rec_{i}"] for i, label in enumerate(meta["labels"])}
    extra = {k: z[k] for k in z.files if k.startswith(("station_", "spatial_"))}
    for u in meta["units"]:
        u["land_covers"] = [tuple(x) for x in u["land_covers"]]
        if u.get("slope"):
            u["slope"] = tuple(u["slope"])
    return meta, forcing, z["outlet"], records, extra


def run_oracle(meta, forcing, labels=None):
    from oracle import hb_oracle

    n = len(next(iter(forcing.values())))
    om = hb_oracle.OracleModel(meta["structures"], meta["units"], meta["solver"], n)
    for k, v in forcing.items():
        om.set_forcing(k, v)
    for label in (meta["labels"] if labels is None else labels):
        om.record(label)
    om.run()
    return om


def same(a, b):
    """Bit-exact including NaN placement."""
    return np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True)
