"""Action dates -> time-step indices. Host-side restatement of TimeMachine::IncrementTime +
ActionsManager::DateUpdate (core/src/base/TimeMachine.cpp:49-59, core/src/actions/ActionsManager.cpp:80-104,
Action.cpp:42-57): after time step k-1 the clock moves to date_k = start + k*dt and DateUpdate(date_k) fires
  * every recursive action whose (month, day) equals date_k's, in the order the actions were added,
  * then every sporadic item with date <= date_k not yet applied, in date order.
So an action "fires before step k" with k >= 1; nothing fires before step 0, and nothing during spin-up.
"""
from __future__ import annotations

import math

import numpy as np

MJD_EPOCH = np.datetime64("1858-11-17", "D")


def step_dates(mjd_start: float, n_steps: int, dt_days: float = 1.0):
    """Calendar (year, month, day) of every step's date."""
    days = (np.floor(mjd_start + np.arange(n_steps) * dt_days)).astype("int64")
    d = MJD_EPOCH + days.astype("timedelta64[D]")
    y = d.astype("datetime64[Y]").astype(int) + 1970
    m = d.astype("datetime64[M]").astype(int) % 12 + 1
    day = (d - d.astype("datetime64[M]")).astype(int) + 1
    return y, m, day


def recursive_steps(mjd_start: float, n_steps: int, month: int, day: int, dt_days: float = 1.0):
    """Steps k >= 1 whose date has the given month and day (Action::AddRecursiveDate clamps the day)."""
    days_in_month = [31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31]
    month = month if 1 <= month <= 12 else 1
    day = max(1, min(day, days_in_month[month - 1]))
    _, m, d = step_dates(mjd_start, n_steps, dt_days)
    k = np.nonzero((m == month) & (d == day))[0]
    return [int(i) for i in k if i >= 1]


def sporadic_step(mjd_start: float, date_mjd: float, dt_days: float = 1.0) -> int:
    """First step k >= 1 with date_k >= date (items dated at or before the start fire after the first step)."""
    return max(1, int(math.ceil((date_mjd - mjd_start) / dt_days - 1e-12)))


def check_fraction(fraction: float) -> float:
    """Action::CheckLandCoverAreaFraction (Action.cpp:71-91)."""
    if abs(fraction) <= 1e-8:
        return 0.0
    if abs(1 - fraction) <= 1e-8:
        return 1.0
    if fraction < 0 or fraction > 1:
        raise ValueError(f"The fraction ({fraction}) is not in the range [0 .. 1]")
    return fraction
