#!/usr/bin/env python
"""Per-SASS-instruction execution counts of one ncu source-page CSV (ncu -i x.ncu-rep --page source --csv), normalised per
warp-step (= executions per 32 HRU*timesteps): python tools/sass_counts.py <source.csv> <P*U*T> [min_count]
Prints: index, executions per warp-step, average active lanes, stall samples, SASS; and the opcode histogram."""
import csv
import sys
from collections import Counter


def load(path, hru_steps):
    rows = list(csv.reader(open(path)))
    hdr_i = next(i for i, r in enumerate(rows) if "Source" in r and "Instructions Executed" in r)
    hdr, data = rows[hdr_i], rows[hdr_i + 1:]
    iS, iE, iT, iSm = (hdr.index(k) for k in ("Source", "Instructions Executed", "Thread Instructions Executed", "# Samples"))
    ws = hru_steps / 32
    out = []
    for k, r in enumerate(data):
        if len(r) <= iT:
            continue
        e = float(r[iE])
        out.append((k, r[iS].strip(), e / ws, float(r[iT]) / max(e, 1), int(r[iSm])))
    return out


def main():
    path, n = sys.argv[1], float(sys.argv[2])
    thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
    rows = load(path, n)
    hist = Counter()
    for k, src, e, lanes, smp in rows:
        op = src.split()[1] if src.startswith("@") else src.split()[0]
        hist[op.split(".")[0]] += e
        if e >= thr:
            print(f"{k:5d} {e:7.3f} {lanes:5.1f} {smp:6d}  {src}")
    tot = sum(hist.values())
    print(f"# total warp instructions per warp-step: {tot:.1f}")
    for op, e in hist.most_common(25):
        print(f"# {op:10s} {e:8.1f} {100 * e / tot:5.1f} %")


if __name__ == "__main__":
    main()
