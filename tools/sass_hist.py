#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of a cubin / .so (no GPU needed).

usage: tools/sass_hist.py <file> [kernel-substring] [--top N]
Groups opcodes by issue pipe as measured in /opt/skills/guides/B300_MICROARCH.md
(fma pipe: IMAD/FFMA..., alu pipe: IADD3/LOP3/SHF/PRMT/SEL/ISETP...).
"""
import collections
import re
import subprocess
import sys

ALU = ("IADD3", "IADD", "LOP3", "SHF", "PRMT", "SEL", "ISETP", "LEA", "FLO", "POPC", "IABS", "IMNMX", "VIADD", "BMSK", "SGXT", "PLOP3", "VIMNMX", "LOP", "MOV", "CS2R", "UMOV")
FMA = ("IMAD", "FFMA", "FMUL", "FADD", "HFMA2", "IDP")
MEM = ("LDG", "STG", "LDS", "STS", "LDC", "LDL", "STL", "ATOM", "RED", "LDSM", "STSM", "ULDC", "LDCU")


def pipe(op):
    base = op.split(".")[0]
    if base in FMA:
        return "fma"
    if base in ALU:
        return "alu"
    if base in MEM:
        return "mem"
    return "other"


def main():
    path = sys.argv[1]
    sub = sys.argv[2] if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else ""
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 12
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout
    cur, hist = None, {}
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            hist[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            hist[cur][m.group(1)] += 1
    for k, h in hist.items():
        if sub not in k:
            continue
        tot = sum(h.values())
        pipes = collections.Counter()
        for op, c in h.items():
            pipes[pipe(op)] += c
        print("== %s  total=%d  %s" % (k, tot, dict(pipes)))
        byb = collections.Counter()
        for op, c in h.items():
            byb[op.split(".")[0] + ("." + op.split(".")[1] if op.startswith("IMAD.") and len(op.split(".")) > 1 else "")] += c
        print("   " + "  ".join("%s:%d" % x for x in byb.most_common(top)))


if __name__ == "__main__":
    main()
