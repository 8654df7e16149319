#!/usr/bin/env python
"""Text summaries of Nsight Compute output for profiles/ (run where `ncu` is installed; no GPU needed to READ a report).

    python benchmarks/ncu_summary.py report  gpurun_out/prof.ncu-rep   > profiles/rN_ncu_<kernel>.txt
    python benchmarks/ncu_summary.py launches gpurun_out/launches.csv  > profiles/rN_launch_list_summary.txt
    python benchmarks/ncu_summary.py traffic gpurun_out/prof.ncu-rep <key> <profiles/summary.txt> [kernel-name-substring]
        records dram__bytes_read.sum + dram__bytes_write.sum per launch of the (matching) kernel in profiles/ncu_traffic.json
        under <key>, with the summary file and the git commit it was captured at — bench.py reads `roofline.traffic` from there
"""
import csv
import io
import subprocess
import sys
from collections import Counter, defaultdict

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "launch__waves_per_multiprocessor",
           "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
           "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_elapsed.max",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"]


def ncu_csv(rep, page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def report(rep):
    rows = ncu_csv(rep, "raw")
    hdr, units, data = rows[0], rows[1], rows[2:]
    names = [r[hdr.index("Kernel Name")] for r in data]
    print("# ncu --set full --clock-control none --import-source on  (report:", rep.split("/")[-1] + ")")
    print("kernel", names)
    for mname in METRICS:
        if mname in hdr:
            k = hdr.index(mname)
            print(mname, f"[{units[k]}]", [r[k] for r in data])
    src = ncu_csv(rep, "source")
    # the source page repeats a header block per kernel; take the first kernel's SASS rows
    h = next(i for i, r in enumerate(src) if r and r[0] == "Address")
    cols = src[h]
    iS, iI, iN = cols.index("Source"), cols.index("Instructions Executed"), cols.index("# Samples")
    ops, samples = Counter(), Counter()
    total = 0
    stall_cols = [k for k, c in enumerate(cols) if c.startswith("stall_")]
    stalls = Counter()
    for r in src[h + 1:]:
        if len(r) <= iI or not r[iI].isdigit():
            if r and r[0] in ("Kernel Name", "Address"):
                break
            continue
        toks = r[iS].split()
        op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
        ops[op] += int(r[iI])
        samples[op] += int(r[iN]) if r[iN].isdigit() else 0
        total += int(r[iI])
        for k in stall_cols:
            if r[k].isdigit():
                stalls[cols[k]] += int(r[k])
    print("total warp instructions executed (first kernel)", total)
    print("opcode  executed  share  stall-samples")
    for op, n in ops.most_common(22):
        print(f"{op:12s} {n:10d}  {100.0 * n / total:4.1f}%  {samples[op]}")
    print("stall samples:", stalls.most_common(10))


def launches(path):
    rows = [r for r in csv.reader(open(path)) if r and r[0].isdigit()]
    agg = defaultdict(lambda: [0, 0.0])
    # columns: ID, Process ID, Process Name, Host Name, Kernel Name, Context, Stream, Block Size, Grid Size, Device, CC, Section, Metric, Unit, Value
    for r in rows:
        if "gpu__time_duration.sum" not in r:
            continue
        name, unit, val = r[4], r[-2], float(r[-1].replace(",", ""))
        us = val / 1000.0 if unit in ("nsecond", "ns") else val * 1000.0 if unit in ("msecond", "ms") else val
        agg[name][0] += 1
        agg[name][1] += us
    tot = sum(v[1] for v in agg.values())
    print("# (cold-cache, serialised launches: compare SHARES, not absolutes)   launches  total_us  share  avg_us  kernel")
    for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{n:5d} {us:10.1f} {100.0 * us / tot:6.1f}% {us / n:9.2f}  {name[:110]}")


def traffic(rep, key, source, match=""):
    import json
    import os
    rows = ncu_csv(rep, "raw")
    hdr, units, data = rows[0], rows[1], rows[2:]
    kn, kr, kw = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    vals = [(float(r[kr].replace(",", "")) * scale[units[kr]], float(r[kw].replace(",", "")) * scale[units[kw]], r[kn]) for r in data if match in r[kn]]
    if not vals:
        raise SystemExit(f"no kernel matching {match!r} in {rep}")
    rd, wr = sum(v[0] for v in vals) / len(vals), sum(v[1] for v in vals) / len(vals)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    path = os.path.join(root, "profiles", "ncu_traffic.json")
    doc = json.load(open(path)) if os.path.exists(path) else {}
    commit = subprocess.run(["git", "-C", root, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    doc[key] = {"bytes": int(rd + wr), "dram_bytes_read": int(rd), "dram_bytes_write": int(wr), "launches_averaged": len(vals),
                "kernel": vals[0][2][:160], "source": source, "commit": commit}
    json.dump(doc, open(path, "w"), indent=1, sort_keys=True)
    print(json.dumps(doc[key]))


if __name__ == "__main__":
    if sys.argv[1] == "traffic":
        traffic(*sys.argv[2:6])
    else:
        {"report": report, "launches": launches}[sys.argv[1]](sys.argv[2])
