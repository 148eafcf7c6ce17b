#!/usr/bin/env python
"""Summarise ncu outputs (read here, no GPU needed) into small text/JSON files under profiles/.

    python tools/ncu_summary.py launches gpurun_out/r01_launches.csv profiles/r01_launches.txt
    python tools/ncu_summary.py report   gpurun_out/r01_scan_stream.ncu-rep profiles/r01_scan_stream.txt [traffic-key]
"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__waves_per_multiprocessor", "sm__cycles_elapsed.max", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
]
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def launches(src, dst):
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else v * 1e3 if r[ui] == "ms" else v
        a = agg.setdefault(r[ki], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised): {src}\n")
        f.write(f"# {len(rows) - 1} launches, {tot:.1f} us total; compare SHARES, not absolutes\n")
        f.write(f"{'us':>12} {'n':>5} {'share':>7}  kernel\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{v[1]:12.1f} {v[0]:5d} {100 * v[1] / tot:6.1f}%  {k[:110]}\n")
    print(open(dst).read())


def report(src, dst, key=None):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = [f"# ncu --set full --clock-control none --import-source on: {src}"]
    traffic = None
    for li, r in enumerate(rows[2:]):
        name = r[hdr.index("Kernel Name")]
        lines.append(f"## launch {li}: {name[:100]}")
        vals = {}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                vals[k] = (r[i], units[i])
                lines.append(f"{k:90s} {r[i]:>18s} {units[i]}")
        if "dram__bytes_read.sum" in vals:
            rd = float(vals["dram__bytes_read.sum"][0].replace(",", "")) * UNIT.get(vals["dram__bytes_read.sum"][1], 1)
            wr = float(vals["dram__bytes_write.sum"][0].replace(",", "")) * UNIT.get(vals["dram__bytes_write.sum"][1], 1)
            traffic = rd + wr
            lines.append(f"{'dram traffic per launch (read + write)':90s} {traffic:18.0f} byte")
    # instruction mix + stall samples from the source page
    src_csv = subprocess.run(["ncu", "-i", src, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = list(csv.reader(io.StringIO(src_csv)))
    hi = [i for i, r in enumerate(srows) if r and r[0] == "Address"]
    if hi:
        h = srows[hi[0]]
        data = srows[hi[0] + 1:(hi[1] - 1 if len(hi) > 1 else len(srows))]
        ie, isrc, ist = h.index("Instructions Executed"), h.index("Source"), h.index("Warp Stall Sampling (All Samples)")
        ops, stall, tot, ts = collections.Counter(), collections.Counter(), 0, 0
        for r in data:
            try:
                n = int(r[ie])
            except ValueError:
                continue
            toks = r[isrc].split()
            op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
            ops[op] += n
            tot += n
            s = int(r[ist] or 0)
            stall[op] += s
            ts += s
        lines.append(f"## SASS: {len(data)} instructions, {tot} warp-instructions executed, {ts} stall samples")
        for k, v in ops.most_common(14):
            lines.append(f"{k:10s} {v:14d} {100 * v / tot:5.1f}% of executed   {100 * stall[k] / max(ts, 1):5.1f}% of stall samples")
    with open(dst, "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines))
    if key and traffic is not None:
        tpath = os.path.join(os.path.dirname(dst), "traffic.json")
        t = json.load(open(tpath)) if os.path.exists(tpath) else {}
        t[key] = traffic
        json.dump(t, open(tpath, "w"), indent=1)


if __name__ == "__main__":
    {"launches": launches, "report": report}[sys.argv[1]](*sys.argv[2:])
