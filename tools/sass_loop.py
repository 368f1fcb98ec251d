#!/usr/bin/env python
"""cuobjdump -sass output (file) -> for each kernel matching a pattern, the instruction mix of its hottest
(largest backward-branch) loop: total, FP64 (DFMA/DMUL/DADD/DSETP/F2F.F64/I2F.F64), MUFU, LDS, LDL/STL."""
import re, sys, collections
pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else re.compile(".")
funcs = collections.OrderedDict(); cur = None
for line in open(sys.argv[1]):
    m = re.search(r"Function : (\S+)", line)
    if m: cur = m.group(1); funcs[cur] = []; continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m and cur: funcs[cur].append((int(m.group(1), 16), m.group(2).strip()))
for name, ins in funcs.items():
    if not pat.search(name): continue
    best = None; bestn = -1
    for addr, text in ins:
        m = re.search(r"BRA\S*\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)", text)
        if m:
            tgt = int(m.group(1), 16)
            if tgt < addr:                     # a loop: keep the one holding the most DFMA (the MH step loop)
                n = sum(1 for a, t in ins if tgt <= a <= addr and "DFMA" in t)
                if n > bestn: best, bestn = (tgt, addr), n
    if not best: continue
    body = [t for a, t in ins if best[0] <= a <= best[1]]
    op = lambda t: re.sub(r"^@!?U?P\d+\s+", "", t).split()[0]
    c = collections.Counter(op(t).split(".")[0] for t in body)
    fp64 = sum(v for k, v in c.items() if k in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX")) + sum(1 for t in body if re.search(r"(F2F|I2F|F2I)\S*F64|\.F64", op(t)))
    print("%s\n  loop 0x%x-0x%x: %d instr, fp64 %d (DFMA %d DMUL %d DADD %d DSETP %d), MUFU %d, LDS %d, LDG %d, STG %d, LDL %d, STL %d, LDC %d, IMAD %d, LOP3 %d, BRA %d" % (
        name, best[0], best[1], len(body), fp64, c["DFMA"], c["DMUL"], c["DADD"], c["DSETP"], c["MUFU"], c["LDS"], c["LDG"], c["STG"], c["LDL"], c["STL"], c["LDC"] + c["LDCU"], c["IMAD"], c["LOP3"], c["BRA"]))
