"""Join an ncu SASS source page (CSV) with nvdisasm -g line info of the same kernel: executed warp instructions
and stall samples per CUDA source line / inlined function."""
import collections
import csv
import re
import sys

sass_file, ncu_csv = sys.argv[1], sys.argv[2]
line_of = []  # per instruction (in order): (file, line, inline chain)
cur = ("?", 0)
for ln in open(sass_file):
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s", ln):
        line_of.append(cur)
rows = list(csv.reader(open(ncu_csv)))
h = rows[1]
ii, sm = h.index("Instructions Executed"), h.index("# Samples")
data = [(int(r[ii]), int(r[sm])) for r in rows[2:] if len(r) > sm]
print("instructions: sass", len(line_of), "ncu", len(data))
n = min(len(line_of), len(data))
agg = collections.defaultdict(lambda: [0, 0])
for i in range(n):
    agg[line_of[i]][0] += data[i][0]
    agg[line_of[i]][1] += data[i][1]
tot = sum(v[0] for v in agg.values())
tots = sum(v[1] for v in agg.values())
src = {}
for f in ("hbk_device.cuh", "hbk_engine.cu"):
    try:
        src[f] = open("/root/repo/hydrobricks_b200/csrc/" + f).read().splitlines()
    except OSError:
        pass
for (f, l), (n_i, n_s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[: int(sys.argv[3]) if len(sys.argv) > 3 else 45]:
    text = src.get(f, [""] * (l + 1))[l - 1].strip() if f in src and l - 1 < len(src[f]) else ""
    print(f"{100 * n_i / tot:5.2f}% inst {100 * n_s / tots:5.2f}% smpl  {f}:{l:<4d} {text[:90]}")
