#!/usr/bin/env python
"""Per-kernel SASS summary of libcodlad_b200.so: the Blackwell-specific instructions that prove the tcgen05 / TMEM / TMA path
(UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG = TMA tensor load / store, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit,
SYNCS = mbarrier ops, MUFU.TANH = the GELU), registers and code size.  Writes profiles/<round>_sass_summary.md.

    python tools/sass_summary.py r02
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "codlad_b200", "libcodlad_b200.so")
OPS = ["UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "STTM", "UTCBAR", "SYNCS", "MUFU.TANH", "MUFU", "HFMA2", "FFMA", "LDG", "STG", "ATOM", "RED", "BAR"]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True, check=True).stdout
    regs = {}
    name = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            name = m.group(1)
        m = re.search(r"REG:(\d+)", line)
        if m and name:
            regs[name] = int(m.group(1))
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    counts, size, cur = {}, collections.Counter(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            size[cur] += 1
            op = m.group(1)
            for o in OPS:
                if op == o or op.startswith(o + "."):
                    counts[cur][o] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
    short = lambda s: re.sub(r"\(.*", "", re.sub(r"cb2::\(anonymous namespace\)::|cb2::|\(anonymous namespace\)::|void ", "", s))
    rows = sorted(zip(demangle, counts), key=lambda kv: -size[kv[1]])
    out = [f"# SASS summary of codlad_b200/libcodlad_b200.so ({tag})", "",
           f"`cuobjdump -sass` / `-res-usage`; cubins: {', '.join(arch)} only.  Counts are static instruction counts per kernel.", "",
           "| kernel | regs | SASS instr | " + " | ".join(OPS) + " |", "|---|---|---|" + "---|" * len(OPS)]
    for dm, k in rows:
        out.append(f"| `{short(dm)}` | {regs.get(k, '')} | {size[k]} | " + " | ".join(str(counts[k][o]) if counts[k][o] else "" for o in OPS) + " |")
    tot = collections.Counter()
    for k in counts:
        tot.update(counts[k])
    out += ["", "Totals: " + ", ".join(f"{o} {tot[o]}" for o in OPS if tot[o]), ""]
    path = os.path.join(ROOT, "profiles", f"{tag}_sass_summary.md")
    with open(path, "w") as f:
        f.write("\n".join(out))
    print(path)


if __name__ == "__main__":
    main()
