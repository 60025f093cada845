"""Summarise an ncu capture (raw + source CSV pages exported with `ncu -i X.ncu-rep --page raw|source --csv`)."""
import collections
import csv
import re
import sys


def main(prefix):
    rows = list(csv.reader(open(prefix + "_source.csv")))
    h = rows[1]
    si, ii, sm = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
    stalls = [i for i, k in enumerate(h) if k.startswith("stall_") and "Not Issued" not in k]
    ops, samp, stallsum = collections.Counter(), collections.Counter(), collections.Counter()
    tot = tots = 0
    for r in rows[2:]:
        if len(r) <= sm:
            continue
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[si].strip())
        op = m.group(2).split(".")[0] if m else r[si][:10]
        n, s = int(r[ii]), int(r[sm])
        ops[op] += n
        tot += n
        samp[op] += s
        tots += s
        for i in stalls:
            stallsum[h[i]] += int(r[i])
    raw = list(csv.reader(open(prefix + "_raw.csv")))
    rh, ru, rv = raw[0], raw[1], raw[2]
    keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
            "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
            "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
            "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum"]
    print("## Raw metrics\n")
    for i, k in enumerate(rh):
        if k in keys:
            print(f"- `{k}` = {rv[i]} {ru[i]}")
    print(f"\n## SASS opcode histogram (share of {tot} warp instructions / of stall samples)\n")
    for op, n in ops.most_common(22):
        print(f"- {op}: {n / tot * 100:.2f} % / {samp[op] / tots * 100:.2f} %")
    print("\n## Warp stall reasons (all samples)\n")
    ts = sum(stallsum.values())
    for k, v in stallsum.most_common(10):
        print(f"- {k}: {v / ts * 100:.2f} %")


if __name__ == "__main__":
    main(sys.argv[1])
