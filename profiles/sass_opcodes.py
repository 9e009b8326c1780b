#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of a built object / shared library (cuobjdump -sass).
Usage: sass_opcodes.py libmod_b200.so [kernel-name-substring ...] > profiles/rNN_sass_opcodes.txt"""
import collections
import re
import subprocess
import sys


def main(path, *filters):
    out = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = kernels.setdefault(name, collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*(?:\.[A-Z0-9_]+)*)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
    for name, ops in kernels.items():
        if filters and not any(f in name for f in filters):
            continue
        short = re.sub(r"\(.*", "", name.replace("(anonymous namespace)::", "")).replace("void ", "").replace("mod::", "")
        total = sum(ops.values())
        base = collections.Counter()
        for op, n in ops.items():
            base[op.split(".")[0]] += n
        print(f"== {short}: {total} SASS instructions")
        print("   by mnemonic: " + ", ".join(f"{k} {v}" for k, v in base.most_common(40)))
        marks = {k: v for k, v in ops.items() if re.match(r"(FFMA2|FADD2|FMUL2|STG\.E\.(ENL2\.)?256|LDG\.E\.128|UBLKCP|UTMALDG|UTMASTG|REDUX|MATCH|UCGABAR|SYNCS|CCTL|ATOM|RED|SHFL|VOTE|BAR)", k)}
        print("   marks: " + ", ".join(f"{k} {v}" for k, v in sorted(marks.items())))


if __name__ == "__main__":
    main(*sys.argv[1:])
