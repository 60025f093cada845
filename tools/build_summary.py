#!/usr/bin/env python
"""Static evidence of one libhbk build (no GPU needed): per kernel of the sm_100a cubin inside the library — registers,
stack frame, static SASS instruction count and the counts of the opcodes that matter here (UBLKCP = cp.async.bulk / TMA,
SYNCS = mbarrier, DFMA / DADD / DMUL / DSETP, LDL / STL = spill traffic) — plus the md5 of the cubin.

    python tools/build_summary.py [hydrobricks_b200/libhbk.so] > profiles/<tag>_build_summary.md
"""
import hashlib
import re
import subprocess
import sys
import tempfile
from collections import Counter
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
OPS = ["UBLKCP", "SYNCS", "DFMA", "DADD", "DMUL", "DSETP", "MUFU", "LDL", "STL", "BAR", "ATOMG", "LDG", "STG"]


def main():
    lib = Path(sys.argv[1]) if len(sys.argv) > 1 else ROOT / "hydrobricks_b200" / "libhbk.so"
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.check_call(["cuobjdump", "-xelf", "all", str(lib.resolve())], cwd=tmp, stdout=subprocess.DEVNULL)
        cubins = sorted(Path(tmp).glob("*.cubin"))
        assert cubins, "no cubin in the library"
        cubin = cubins[0]
        md5 = hashlib.md5(cubin.read_bytes()).hexdigest()
        arch = re.search(r"sm_\d+a?", cubin.name).group(0)
        usage = subprocess.run(["cuobjdump", "-res-usage", str(cubin)], capture_output=True, text=True).stdout
        sass = subprocess.run(["cuobjdump", "-sass", str(cubin)], capture_output=True, text=True).stdout
        demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n  # noqa: E731
    res = {}
    name = None
    for line in usage.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            name = m.group(1)
            continue
        m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+)", line)
        if m and name:
            res[name] = tuple(int(x) for x in m.groups())
    counts, cur = {}, None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = counts.setdefault(m.group(1), Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur is not None:
            cur["_total"] += 1
            cur[m.group(1)] += 1
    print(f"# Build summary of `{lib.name}`: {arch} cubin md5 `{md5}`, {len(res)} kernels / device functions\n")
    print("Flags: `-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false` (`__graft_entry__.build`). Static counts "
          "(`cuobjdump -res-usage`, `-sass`); `hbkRunUnits<SOLVER, NSOIL, GLACIER, DAILY, SHAPE>`: SOLVER 0 Euler, 1 Heun, 2 RK4, "
          "3 sequential family; SHAPE 1 = ensemble (32 sets x 4 units), 2 = basin (1 set x 128 units), 0 = generic.\n")
    print("| kernel | registers | stack B | SASS instr. | " + " | ".join(OPS) + " |")
    print("|---|---|---|---|" + "---|" * len(OPS))
    for mangled in sorted(res, key=lambda n: demangle(n)):
        c = counts.get(mangled, Counter())
        reg, stack, _ = res[mangled]
        nice = demangle(mangled).replace("hbk::", "").replace("(hbk::KArgs)", "").replace("(KArgs)", "")
        nice = re.sub(r"^void ", "", nice)[:80]
        print(f"| `{nice}` | {reg} | {stack} | {c['_total']} | " + " | ".join(str(c[o]) for o in OPS) + " |")


if __name__ == "__main__":
    main()
