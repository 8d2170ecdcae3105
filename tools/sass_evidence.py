#!/usr/bin/env python
"""Developer tool (no GPU needed): static SASS evidence of the shipped library, for profiles/.

  python tools/sass_evidence.py > profiles/rNN_sass_opcodes.txt

For every kernel of seq2squiggle_b200/libs2s_b200.so (cuobjdump -sass): instruction count and the static counts of the
opcodes that identify the execution path — UTC*MMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG (TMA
tensor loads / stores), SYNCS (mbarrier), MUFU.EX2, F2FP, legacy HMMA (must be absent) — next to the general mix."""
import os
import re
import subprocess
import sys
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "seq2squiggle_b200", "libs2s_b200.so")
KEY = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "MUFU.EX2", "F2FP", "HMMA", "LDS",
       "STS", "LDG", "STG", "LDC", "LDCU", "BAR"]


def strip_params(name: str) -> str:
    """``f<a, b>(args)`` -> ``f<a, b>``: cut at the first '(' outside the template argument list."""
    depth = 0
    for i, ch in enumerate(name):
        depth += (ch == "<") - (ch == ">")
        if ch == "(" and depth == 0:
            return name[:i]
    return name


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else LIB
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    regs = {}
    res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
    for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+)[^\n]*SHARED:(\d+)", res):
        regs[m.group(1)] = (int(m.group(2)), int(m.group(3)))
    kernels, cur = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*(?:\.[A-Z0-9_]+)*)", line)
        if m and cur is not None:
            op = m.group(1)
            cur["_total"] += 1
            parts = op.split(".")
            cur[parts[0]] += 1
            if parts[0] == "MUFU" and len(parts) > 1:
                cur["MUFU." + parts[1]] += 1
    demangle = subprocess.run(["cu++filt"] + list(kernels), capture_output=True, text=True).stdout.splitlines()
    names = dict(zip(kernels, demangle)) if len(demangle) == len(kernels) else {k: k for k in kernels}
    print(f"# {os.path.relpath(lib, ROOT)}: static SASS opcode counts per kernel (cuobjdump -sass, sm_100a)")
    print(f"# {'kernel':58s} {'instr':>6s} {'regs':>4s}  " + " ".join(f"{k:>8s}" for k in KEY))
    for k, c in sorted(kernels.items(), key=lambda kv: -kv[1]["_total"]):
        short = strip_params(names[k].replace("s2s::(anonymous namespace)::", "").replace("s2s::<unnamed>::", "")
                             .replace("s2s::", "").replace("(bool)", ""))
        short = re.sub(r"^void ", "", short)
        r = regs.get(k, ("", ""))[0]
        print(f"{short[:60]:60s} {c['_total']:6d} {str(r):>4s}  " + " ".join(f"{c.get(x, 0):8d}" for x in KEY))
    legacy = sum(c.get("HMMA", 0) for c in kernels.values())
    print(f"# legacy tensor path (HMMA) instructions in the library: {legacy}")


if __name__ == "__main__":
    main()
