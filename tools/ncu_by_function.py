"""Executed warp instructions of one kernel by source FUNCTION: joins the ncu SASS source page (CSV) with `nvdisasm -g`
line info of the same kernel (as tools/ncu_lanes_by_line.py) and maps lines to the device functions of the sources.
usage: ncu_by_function.py kernel.sass source_page.csv hru_timesteps_of_the_launch"""
import collections
import csv
import re
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def function_starts(path):
    out, pending = [], None
    for i, line in enumerate(open(path).read().splitlines(), 1):
        if line.startswith("template <"):
            pending = i
            continue
        m = re.match(r"^(?:__device__|__global__|extern \"C\"|static|inline)[^;]*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\($", line) or \
            re.match(r"^(?:__device__|__global__)[^;(]*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", line)
        if m:
            out.append((pending or i, m.group(1)))
        if not line.startswith(("__", "template", " ")) or m:
            pending = None
    return out


def main(sass, page, hru_steps):
    starts = {"hbk_device.cuh": function_starts(ROOT / "hydrobricks_b200/csrc/hbk_device.cuh"),
              "hbk_engine.cu": function_starts(ROOT / "hydrobricks_b200/csrc/hbk_engine.cu")}
    starts["hbk_engine.cu"] = [(s, n if n != "__launch_bounds__" else "hbkRunUnits (kernel glue)") for s, n in starts["hbk_engine.cu"]]

    def func_of(f, l):
        name = "?"
        for start, n in starts.get(f, []):
            if start <= l:
                name = n
            else:
                break
        return name

    line_of, cur = [], ("?", 0)
    for ln in open(sass):
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s", ln):
            line_of.append(cur)
    rows = list(csv.reader(open(page)))
    h = rows[1]
    ii, ti, sm = h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("# Samples")
    agg = collections.defaultdict(lambda: [0, 0, 0])
    for i, r in enumerate(rows[2:]):
        if len(r) <= ti or i >= len(line_of):
            continue
        f, l = line_of[i]
        key = func_of(f, l) if f in starts else f
        agg[key][0] += int(r[ii])
        agg[key][1] += int(r[ti])
        agg[key][2] += int(r[sm])
    tot = sum(v[0] for v in agg.values())
    tots = sum(v[2] for v in agg.values())
    print("| function | warp instructions | stall samples | lane efficiency | thread instructions per HRU*timestep |")
    print("|---|---|---|---|---|")
    for fn, (ni, nt, ns) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        if ni / tot < 0.002:
            continue
        print(f"| `{fn}` | {100 * ni / tot:.1f} % | {100 * ns / tots:.1f} % | {nt / (32 * ni):.2f} | {nt / hru_steps:.0f} |")
    print(f"\ntotal: {sum(v[1] for v in agg.values()) / hru_steps:.0f} thread instructions per HRU*timestep")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], float(sys.argv[3]))
