#!/usr/bin/env python
"""Reads a HBK_BLOCK_TIMES dump (start ns, end ns, SM per block of the unit launch) and prints the schedule's shape:
block durations by position in the grid, busy slots over time, the tail.   python tools/block_times.py <file> [slots]"""
import sys

import numpy as np

a = np.fromfile(sys.argv[1], dtype=np.uint64).reshape(-1, 3)
slots = int(sys.argv[2]) if len(sys.argv) > 2 else 888
t0 = a[:, 0].min()
start, end = (a[:, 0] - t0) / 1e6, (a[:, 1] - t0) / 1e6
dur = end - start
n = len(a)
print(f"blocks {n}, makespan {end.max():.1f} ms, sum of durations / slots = {dur.sum() / slots:.1f} ms "
      f"-> schedule efficiency {dur.sum() / slots / end.max():.3f}")
print("duration ms: min %.1f  p10 %.1f  median %.1f  p90 %.1f  max %.1f" % (dur.min(), *np.percentile(dur, [10, 50, 90]), dur.max()))
k = 10
for i in range(k):
    sel = slice(i * n // k, (i + 1) * n // k)
    print(f"grid decile {i}: mean duration {dur[sel].mean():7.1f} ms, mean start {start[sel].mean():7.1f} ms")
grid = np.linspace(0, end.max(), 21)
for lo, hi in zip(grid[:-1], grid[1:]):
    mid = (lo + hi) / 2
    busy = int(((start <= mid) & (end > mid)).sum())
    print(f"t = {mid:7.1f} ms: {busy:4d} blocks resident")
