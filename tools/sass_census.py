#!/usr/bin/env python
"""SASS census of the built library (no GPU needed): which kernels issue tcgen05 / TMA / tensor-memory instructions, how many.

    python tools/sass_census.py [libclipk.so] > profiles/rNN_sass_census.md

Mnemonics (B200_PROFILING.md): UTCHMMA = tcgen05.mma (kind::f16), UTMALDG / UTMASTG = TMA bulk tensor load / store, LDTM / STTM =
tcgen05.ld / tcgen05.st (tensor memory <-> registers), UTCBAR = tcgen05.commit -> mbarrier, SYNCS = mbarrier ops, USETMAXREG = setmaxnreg,
ELECT = elect.sync."""
import collections
import os
import re
import subprocess
import sys

KEYS = ["UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "SYNCS", "USETMAXREG", "ELECT", "MUFU.EX2", "MUFU.TANH", "REDG", "ATOMG"]


def census(lib):
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    cur, stats = None, collections.OrderedDict()
    for ln in sass.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1); stats[cur] = collections.Counter(); continue
        if cur is None:
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if m:
            stats[cur]["_n"] += 1
            for k in KEYS:
                if m.group(1).startswith(k):
                    stats[cur][k] += 1
    return stats


def main():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(root, "easynlp_b200", "lib", "libclipk.so")
    stats = census(lib)
    names = subprocess.run(["c++filt"], input="\n".join(stats), capture_output=True, text=True).stdout.splitlines()
    fam = collections.OrderedDict()
    for mangled, name in zip(stats, names):
        c = stats[mangled]
        short = re.sub(r"\(.*", "", name.replace("(anonymous namespace)::", "")).replace("clipk::", "").replace("void ", "")
        family = re.sub(r"<.*", "", short)
        fam.setdefault(family, []).append((short, c))
    stamp = os.path.join(os.path.dirname(lib), "libclipk.sha256")
    print(f"# SASS census of {os.path.relpath(lib, root)} (source digest {open(stamp).read().strip()[:16] if os.path.exists(stamp) else '?'})\n")
    print(f"{len(stats)} device functions.  Per kernel family: instances, SASS instructions of the largest instance, and the count of each\n"
          "tensor-core / TMA / tensor-memory mnemonic in that instance (`cuobjdump -sass`, produced by tools/sass_census.py without a GPU).\n")
    print("| kernel family | instances | SASS instrs | " + " | ".join(KEYS) + " |")
    print("|---|---|---|" + "---|" * len(KEYS))
    for family, items in fam.items():
        short, c = max(items, key=lambda it: it[1]["_n"])
        if not any(c[k] for k in ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM")):
            continue
        print(f"| `{short}` | {len(items)} | {c['_n']} | " + " | ".join(str(c[k]) if c[k] else "" for k in KEYS) + " |")
    plain = [f for f, items in fam.items() if not any(c[k] for _, c in items for k in ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM"))]
    print(f"\n{len(plain)} families without tensor-core / TMA instructions (elementwise, reductions, optimizer, loss rows, preprocessing, peer-memory):\n"
          + ", ".join(f"`{p}`" for p in plain))


if __name__ == "__main__":
    main()
