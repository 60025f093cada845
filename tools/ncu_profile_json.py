#!/usr/bin/env python
"""profiles/kernel_profile.json from one ncu capture of the hot kernel (what bench.py prints beside the roofline):

    python tools/ncu_profile_json.py <version> <raw.csv> <source.csv> <hru_steps of the captured launch> [traffic.csv]

raw.csv / source.csv: `ncu -i x.ncu-rep --page raw|source --csv` of a `--set full` capture (tools/gpu_kbench.sh KB_NCU=…);
traffic.csv: `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --csv` of the FULL-SIZE C2 launch
(bench.py --steps 1 --warmup 0), whose bytes are what `roofline.traffic` reports.
"""
import csv
import json
import sys
from collections import Counter
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def raw_metrics(path):
    rows = list(csv.reader(open(path)))
    hdr, units, vals = rows[0], rows[1], rows[-1]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "usecond": 1e-3, "msecond": 1, "second": 1e3,
             "nsecond": 1e-6}
    out = {}
    for k, u, v in zip(hdr, units, vals):
        if (k.startswith("dram__bytes") or k == "gpu__time_duration.sum") and u in scale:
            try:
                v = str(float(v.replace(",", "")) * scale[u])
            except ValueError:
                pass
        out[k] = v
    return out


def opcode_thread_instructions(path):
    rows = list(csv.reader(open(path)))
    i = next(k for k, r in enumerate(rows) if "Source" in r and "Thread Instructions Executed" in r)
    hdr = rows[i]
    iS, iT, iE = hdr.index("Source"), hdr.index("Thread Instructions Executed"), hdr.index("Instructions Executed")
    thread, warp = Counter(), Counter()
    for r in rows[i + 1:]:
        if len(r) <= iT:
            continue
        src = r[iS].strip()
        op = (src.split()[1] if src.startswith("@") else src.split()[0]).split(".")[0]
        thread[op] += float(r[iT])
        warp[op] += float(r[iE])
    return thread, warp


def main():
    version, raw, source, n = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
    m = raw_metrics(raw)
    thread, warp = opcode_thread_instructions(source)
    f = lambda k: float(m[k]) if k in m and m[k] not in ("", "n/a") else None  # noqa: E731
    fp64_ops = {k: thread[k] / n for k in ("DADD", "DMUL", "DFMA", "DSETP")}
    out = {
        "version": version,
        "kernel": m.get("Kernel Name"),
        "captured_launch_hru_steps": n,
        "gpu_time_ms": f("gpu__time_duration.sum"),
        "registers_per_thread": f("launch__registers_per_thread"),
        "fp64_pipe_pct": f("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
        "issue_active_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "eligible_warps_per_cycle": f("smsp__warps_eligible.avg.per_cycle_active"),
        "lanes_per_warp_instruction": f("smsp__thread_inst_executed_per_inst_executed.ratio"),
        "warp_instructions": f("smsp__inst_executed.sum"),
        "thread_instructions_per_hru_step": sum(thread.values()) / n,
        "fp64_thread_instructions_per_hru_step": fp64_ops,
        "flops_executed_per_hru_step": fp64_ops["DADD"] + fp64_ops["DMUL"] + 2 * fp64_ops["DFMA"],
        "dram_bytes_captured_launch": (f("dram__bytes_read.sum") or 0) + (f("dram__bytes_write.sum") or 0),
        "opcode_share_of_warp_instructions_pct": {k: round(100 * v / sum(warp.values()), 2) for k, v in warp.most_common(16)},
    }
    if len(sys.argv) > 5:
        rows = list(csv.reader(l for l in open(sys.argv[5]) if l.startswith('"')))
        hdr = rows[0]
        iK, iN, iV, iU = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
        tot = 0.0
        for r in rows[1:]:
            if "hbkRunUnits" in r[iK] and r[iN].startswith("dram__bytes"):
                tot += float(r[iV].replace(",", "")) * scale.get(r[iU], 1)
        out["dram_bytes_full_c2_launch"] = tot
    (ROOT / "profiles" / "kernel_profile.json").write_text(json.dumps(out, indent=1))
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
