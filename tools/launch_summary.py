#!/usr/bin/env python
"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv`) of ONE training
step into markdown + a small json (per-kernel share of the step; DRAM bytes per GEMM launch = bench.py's roofline.traffic).
usage: tools/launch_summary.py launches.csv out.md out.json [skip_launches] [count_launches]
       tools/launch_summary.py launches.csv out.md out.json step K     (the K-th optimizer step found in the list, 0-based)"""
import csv, json, re, sys
from collections import OrderedDict


def short(name):
    name = re.sub(r"^void ", "", name)
    name = name.replace("clipk::", "")
    m = re.match(r"([A-Za-z0-9_]+)(<[^(]*>)?\(", name)
    if not m:
        return name[:60]
    t = m.group(2) or ""
    return m.group(1) + re.sub(r"\(int\)|\(bool\)", "", t)


def main():
    src, out_md, out_json = sys.argv[1:4]
    step_k = int(sys.argv[5]) if len(sys.argv) > 5 and sys.argv[4] == "step" else None
    skip = int(sys.argv[4]) if len(sys.argv) > 4 and step_k is None else 0
    count = int(sys.argv[5]) if len(sys.argv) > 5 and step_k is None else None
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr, rows = rows[0], rows[1:]
    ix = {h: i for i, h in enumerate(hdr)}
    launches = OrderedDict()
    for r in rows:
        lid = int(r[ix["ID"]])
        d = launches.setdefault(lid, {"name": r[ix["Kernel Name"]]})
        v = float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]]
        if unit in ("us", "usecond"): v *= 1e3
        if unit in ("ms", "msecond"): v *= 1e6
        if unit == "Kbyte": v *= 1e3
        if unit == "Mbyte": v *= 1e6
        if unit == "Gbyte": v *= 1e9
        d[r[ix["Metric Name"]]] = v
    ids = sorted(launches)[skip:]
    if count is not None:
        ids = ids[:count]
    if step_k is not None:
        # a step ends with its last adamw_kernel launch (two parameter groups); it starts right after the previous step's
        ends = [i for i in ids if "adamw_kernel" in launches[i]["name"] and not (i + 1 in launches and "adamw_kernel" in launches[i + 1]["name"])]
        lo = ends[step_k - 1] + 1 if step_k > 0 else ids[0]
        ids = [i for i in ids if lo <= i <= ends[step_k]]
    agg = OrderedDict()
    for i in ids:
        d = launches[i]
        a = agg.setdefault(short(d["name"]), {"n": 0, "ns": 0.0, "rd": 0.0, "wr": 0.0})
        a["n"] += 1; a["ns"] += d.get("gpu__time_duration.sum", 0.0)
        a["rd"] += d.get("dram__bytes_read.sum", 0.0); a["wr"] += d.get("dram__bytes_write.sum", 0.0)
    tot = sum(a["ns"] for a in agg.values()) or 1.0
    fam = lambda k: k.split("<")[0]
    fams = OrderedDict()
    for k, a in agg.items():
        f = fams.setdefault(fam(k), {"n": 0, "ns": 0.0, "rd": 0.0, "wr": 0.0})
        for q in f: f[q] += a[q]
    with open(out_md, "w") as f:
        f.write(f"# ncu launch list of one training step ({len(ids)} launches; serialised, cold-cache per-launch times: the SHARES are what carries over)\n\n")
        f.write("| kernel family | launches | sum ms | share | DRAM GB read | DRAM GB written |\n|---|---|---|---|---|---|\n")
        for k, a in sorted(fams.items(), key=lambda kv: -kv[1]["ns"]):
            f.write(f"| `{k}` | {a['n']} | {a['ns']/1e6:.3f} | {100*a['ns']/tot:.1f} % | {a['rd']/1e9:.2f} | {a['wr']/1e9:.2f} |\n")
        f.write(f"| **total** | {len(ids)} | {tot/1e6:.3f} | | {sum(a['rd'] for a in fams.values())/1e9:.2f} | {sum(a['wr'] for a in fams.values())/1e9:.2f} |\n\n")
        f.write("## by template instance\n\n| kernel | launches | sum ms | share | avg us |\n|---|---|---|---|---|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["ns"]):
            f.write(f"| `{k}` | {a['n']} | {a['ns']/1e6:.3f} | {100*a['ns']/tot:.1f} % | {a['ns']/a['n']/1e3:.1f} |\n")
    g = fams.get("gemm_bf16_kernel", {"n": 0, "ns": 0, "rd": 0, "wr": 0})
    js = {"launches": len(ids), "sum_ms": tot / 1e6, "gemm": {"launches": g["n"], "share_of_step": g["ns"] / tot,
          "dram_bytes_per_launch": (g["rd"] + g["wr"]) / g["n"] if g["n"] else None, "avg_launch_us": g["ns"] / g["n"] / 1e3 if g["n"] else None},
          "shares": {k: a["ns"] / tot for k, a in fams.items()}}
    # the build the list was captured with: bench.py refuses a profile whose hash differs from the loaded library.  The GPU command copies
    # easynlp_b200/lib/libclipk.sha256 next to the csv (<csv>.sha256); fall back to the in-tree stamp.
    import os
    for cand in (src + ".sha256", os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "easynlp_b200", "lib", "libclipk.sha256")):
        if os.path.exists(cand):
            js["libclipk_sha256"] = open(cand).read().strip()
            break
    json.dump(js, open(out_json, "w"), indent=1)
    print("wrote", out_md, out_json, js["gemm"])


if __name__ == "__main__":
    main()
