#!/usr/bin/env python
"""Summarise .ncu-rep files (read here without a GPU) into markdown for profiles/.
usage: tools/ncu_summary.py out.md title rep1.ncu-rep [rep2 ...]"""
import csv, io, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def source_top(rep, n=12):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    kern, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}; kern.append(cur); continue
        if cur is not None:
            cur["rows"].append(r)
    res = []
    for k in kern:
        hdr = k["rows"][0]; idx = {h: i for i, h in enumerate(hdr)}
        data = [r for r in k["rows"][1:] if len(r) > idx["# Samples"] and r[idx["# Samples"]].isdigit()]
        tot = sum(int(r[idx["# Samples"]]) for r in data) or 1
        top = sorted(data, key=lambda r: -int(r[idx["# Samples"]]))[:n]
        res.append((k["name"], tot, [(int(r[idx["# Samples"]]), r[idx["Source"]].strip()) for r in top]))
    return res


def main():
    out, title, reps = sys.argv[1], sys.argv[2], sys.argv[3:]
    with open(out, "w") as f:
        f.write(f"# {title}\n\n")
        for rep in reps:
            hdr, units, rows = raw(rep)
            idx = {h: i for i, h in enumerate(hdr)}
            stalls = [h for h in hdr if h.startswith("smsp__pcsamp_warps_issue_stalled") and not h.endswith("not_issued")]
            src = source_top(rep)
            f.write(f"## {rep.split('/')[-1]}\n\n")
            for ki, r in enumerate(rows):
                name = r[idx["Kernel Name"]]
                f.write(f"### launch {ki}: `{name[:110]}`  grid {r[idx['Grid Size']]} block {r[idx['Block Size']]}\n\n| metric | value | unit |\n|---|---|---|\n")
                for k in KEYS:
                    if k in idx:
                        f.write(f"| {k} | {r[idx[k]]} | {units[idx[k]]} |\n")
                st = sorted(((int(float(r[idx[s]] or 0)), s.replace("smsp__pcsamp_warps_issue_stalled_", "")) for s in stalls), reverse=True)[:6]
                f.write("| top stall reasons (pc samples) | " + ", ".join(f"{n} {c}" for c, n in st) + " | |\n\n")
                if ki < len(src):
                    f.write("hot SASS (samples, instruction):\n\n```\n" + "\n".join(f"{c:6d}  {s[:120]}" for c, s in src[ki][2]) + "\n```\n\n")
    print("wrote", out)


if __name__ == "__main__":
    main()
