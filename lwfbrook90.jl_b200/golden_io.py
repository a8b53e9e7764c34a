"""CSV tables in the layouts of the reference's regression tests.

The reference stores its golden results as CSV written by CSV.jl from DataFrames
(test/03-regression-tests.jl:35-62 ``get_some_states_to_compare``, :231-273 ``get_daily_soilFluxes``; files
``<case>_u_sol_referenceu_mm.csv``, ``..._u_δ.csv``, ``..._u_belowground.csv``, ``<case>_fluxes_referencedf.csv``).
These functions produce and read the same files, so that results of this backend can be diffed against the
reference's goldens — and the reference's own regression tests can read files written here — column for column.

Number format: the shortest string that round-trips the double (what Julia prints), ``NaN`` for NaN, ISO timestamps
with one fractional digit (``2010-01-01T00:00:00.0``) in the ``time`` column of the flux table.
"""
from __future__ import annotations

import csv
import datetime as _dt
import os
from typing import Dict, List

import numpy as np

from . import api


def _fmt(v) -> str:
    if isinstance(v, _dt.datetime):
        return v.strftime("%Y-%m-%dT%H:%M:%S") + f".{v.microsecond // 100000}"
    x = float(v)
    if x != x:
        return "NaN"
    if x in (float("inf"), float("-inf")):
        return "Inf" if x > 0 else "-Inf"
    return repr(x)


def write_table(path: str, table: Dict[str, object]) -> None:
    """One CSV column per key, in insertion order."""
    cols = list(table)
    n = len(table[cols[0]])
    if any(len(table[c]) != n for c in cols):
        raise ValueError("write_table: columns of different lengths")
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "w", encoding="utf-8", newline="") as f:
        w = csv.writer(f, lineterminator="\n")
        w.writerow(cols)
        for i in range(n):
            w.writerow([_fmt(table[c][i]) for c in cols])


def read_table(path: str) -> Dict[str, np.ndarray]:
    """Columns of a golden CSV: float arrays (NaN for ``NaN`` / ``missing``), timestamps as datetime objects."""
    with open(path, encoding="utf-8") as f:
        rows = list(csv.reader(f))
    out: Dict[str, np.ndarray] = {}
    for j, name in enumerate(rows[0]):
        raw = [r[j] for r in rows[1:]]
        try:
            out[name] = np.array([float("nan") if x in ("NaN", "missing", "") else float(x) for x in raw])
        except ValueError:
            out[name] = np.array([_dt.datetime.strptime(x[:19], "%Y-%m-%dT%H:%M:%S") for x in raw], dtype=object)
    return out


def get_some_states_to_compare(sim):
    """(u_mm, u_δ, u_belowground) every 30 days, as test/03-regression-tests.jl:35-62 builds them."""
    sol = sim.solution
    t_out = np.arange(sol.t.min(), sol.t.max() + 1e-9, 30.0)
    ix = sol.ix
    recs = np.stack([sol.record_at(float(t)) for t in t_out])
    u_mm = {"time": t_out.copy()}
    for name, col in (("GWAT", ix.GWAT_mm), ("INTS", ix.INTS_mm), ("INTR", ix.INTR_mm), ("SNOW", ix.SNOW_mm), ("CC", ix.CC),
                      ("SNOWLQ", ix.SNOWLQ), ("RWU", ix.RWU_mm), ("XYLEM", ix.XYLEM_mm)):
        u_mm[name] = recs[:, col]
    u_delta = api.get_delta(sim, days_to_read_out_d=t_out)
    below = api.get_soil_(["SWATI", "W", "ψ", "θ", "K", "δ18O", "δ2H"], sim, days_to_read_out_d=t_out)
    u_below = {k: v for k, v in below.items() if not (k.startswith("K_Lay") or k.startswith("W_Lay"))}
    return u_mm, u_delta, u_below


def write_regression_goldens(sim, folder: str, case: str) -> List[str]:
    """Writes the four golden tables of a simulated case; returns the paths."""
    u_mm, u_delta, u_below = get_some_states_to_compare(sim)
    paths = [os.path.join(folder, f"{case}_u_sol_referenceu_mm.csv"), os.path.join(folder, f"{case}_u_sol_referenceu_δ.csv"),
             os.path.join(folder, f"{case}_u_sol_referenceu_belowground.csv"), os.path.join(folder, f"{case}_fluxes_referencedf.csv")]
    write_table(paths[0], u_mm)
    write_table(paths[1], u_delta)
    write_table(paths[2], u_below)
    write_table(paths[3], api.get_daily_soilFluxes(sim))
    return paths
