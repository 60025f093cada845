"""Lane efficiency per CUDA source line: joins the ncu SASS source page (CSV) with `nvdisasm -g` line info of the same
kernel and ranks the lines by wasted lane slots (32 x warp instructions - thread instructions).
usage: ncu_lanes_by_line.py kernel.sass source_page.csv [n]"""
import collections
import csv
import re
import sys

line_of, cur = [], ("?", 0)
for ln in open(sys.argv[1]):
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s", ln):
        line_of.append(cur)
rows = list(csv.reader(open(sys.argv[2])))
h = rows[1]
ii, ti, sm = h.index("Instructions Executed"), h.index("Thread Instructions Executed"), h.index("# Samples")
data = [(int(r[ii]), int(r[ti]), int(r[sm])) for r in rows[2:] if len(r) > ti]
print("instructions: sass", len(line_of), "ncu", len(data))
agg = collections.defaultdict(lambda: [0, 0, 0])
for i in range(min(len(line_of), len(data))):
    for k in range(3):
        agg[line_of[i]][k] += data[i][k]
tot = sum(v[0] for v in agg.values())
tott = sum(v[1] for v in agg.values())
tots = sum(v[2] for v in agg.values())
print(f"lane efficiency {tott / (32 * tot):.3f}; warp instructions {tot}")
src = {}
for f in ("hbk_device.cuh", "hbk_engine.cu"):
    try:
        src[f] = open("/root/repo/hydrobricks_b200/csrc/" + f).read().splitlines()
    except OSError:
        pass
waste = 32 * tot - tott
key = (lambda kv: -(32 * kv[1][0] - kv[1][1])) if (len(sys.argv) <= 4 or sys.argv[4] == "waste") else (lambda kv: -kv[1][0])
for (f, l), (ni, nt, ns) in sorted(agg.items(), key=key)[: int(sys.argv[3]) if len(sys.argv) > 3 else 40]:
    text = src[f][l - 1].strip() if f in src and l - 1 < len(src[f]) else ""
    print(f"{100 * (32 * ni - nt) / waste:5.2f}% waste {100 * ni / tot:5.2f}% inst {100 * ns / tots:5.2f}% smpl eff {nt / (32 * ni) if ni else 0:4.2f}  {f}:{l:<4d} {text[:70]}")
