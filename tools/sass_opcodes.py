#!/usr/bin/env python
"""SASS opcode histogram of every kernel in libeps_b200.so -> profiles/sass_opcodes.md
(cuobjdump -sass; the Blackwell-native mnemonics are listed first: UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st,
UTCBAR = tcgen05.commit, UTMALDG/UTMASTG/UBLKCP = TMA, SYNCS = mbarrier, REDG/ATOMG = global reductions)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "edge_proposal_sets_b200", "libeps_b200.so")
KEY = ("UTCHMMA", "UTCQMMA", "UTCBAR", "UTCCP", "LDTM", "STTM", "UTCATOM", "UTMALDG", "UTMASTG", "UBLKCP", "UTMAPF",
       "SYNCS", "HMMA", "HMUL2", "F2FP", "FFMA", "FFMA2", "FADD2", "LDG", "STG", "LDS", "STS", "LDGSTS", "REDG", "ATOMG",
       "ATOMS", "RED", "MATCH", "VOTE", "SHFL", "POPC", "CCTL", "BAR", "ELECT")


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\(.*", "", cur).replace("eps::", "")
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*(?:\.[A-Z0-9_.]+)?)", line)
        if m and cur:
            kernels[cur][m.group(1)] += 1
    lines = ["# SASS opcode histogram of `libeps_b200.so` (sm_100a)", "",
             f"`python tools/sass_opcodes.py` — `cuobjdump -sass` of the in-tree library, one row per kernel; counts are static "
             "instruction counts.  Columns: total instructions, then the opcodes that identify the hardware path "
             "(base mnemonic, all modifiers folded), then the full mnemonics of the tensor / TMEM / TMA / mbarrier instructions.", "",
             "| kernel | instrs | " + " | ".join(KEY) + " | tensor / TMEM / TMA / barrier mnemonics |", "|---|---|" + "---|" * (len(KEY) + 1)]
    for name, c in kernels.items():
        base = collections.Counter()
        for op, k in c.items():
            base[op.split(".")[0]] += k
        special = sorted((op, k) for op, k in c.items() if op.split(".")[0] in ("UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG",
                                                                               "UTMASTG", "UBLKCP", "UTCCP", "UTCATOM", "SYNCS"))
        lines.append(f"| `{name}` | {sum(c.values())} | " + " | ".join(str(base.get(k, 0) or "") for k in KEY) + " | " +
                     ", ".join(f"{op} x{k}" for op, k in special) + " |")
    path = os.path.join(ROOT, "profiles", "sass_opcodes.md")
    open(path, "w").write("\n".join(lines) + "\n")
    print(path, len(kernels), "kernels")


if __name__ == "__main__":
    main()
