#!/usr/bin/env python
"""nvcc -Xptxas -v output (stdin) -> one line per kernel: registers, spills, stack, shared memory."""
import re, sys, subprocess
cur = None; rows = []
for line in sys.stdin:
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        cur = {"name": m.group(1)}; rows.append(cur); continue
    if cur is None: continue
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m: cur["stack"], cur["spst"], cur["spld"] = map(int, m.groups())
    m = re.search(r"Used (\d+) registers", line)
    if m: cur["regs"] = int(m.group(1))
names = subprocess.run(["c++filt"], input="\n".join(r["name"] for r in rows), capture_output=True, text=True).stdout.splitlines()
for r, n in zip(rows, names):
    n = re.sub(r"\(.*\)$", "", n).replace("void mcgpu::", "")
    print("%-70s regs %3d  stack %4d  spill st/ld %4d/%4d" % (n, r.get("regs", -1), r.get("stack", 0), r.get("spst", 0), r.get("spld", 0)))
