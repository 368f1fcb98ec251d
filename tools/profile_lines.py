#!/usr/bin/env python
"""Per-source-line instruction counts of one captured launch of an ncu report (needs -lineinfo and
--import-source on):  python tools/profile_lines.py rep.ncu-rep LAUNCH_INDEX [warps*steps divisor] [top N]"""
import csv, subprocess, sys
rep, k = sys.argv[1], int(sys.argv[2])
div = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
top = int(sys.argv[4]) if len(sys.argv) > 4 else 60
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--launch-skip", str(k),
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fname = None; hdr = None; acc = []
for r in rows:
    if not r: continue
    if r[0] in ("File Name", "File Path"): fname = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if r[0] == "Kernel Name": print("#", r[1][:110]); continue
    if hdr and r[0].isdigit():
        ie = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples")
        try: acc.append((float(r[ie]), float(r[isamp] or 0), fname, int(r[0]), r[1].strip()))
        except ValueError: pass
tot = sum(a[0] for a in acc); ts = sum(a[1] for a in acc)
print("# total warp instructions %.0f (%.1f per unit), samples %.0f" % (tot, tot / div, ts))
for n, s, f, ln, src in sorted(acc, reverse=True)[:top]:
    print("%9.1f %5.1f%% smp %5.1f%%  %s:%d  %s" % (n / div, 100 * n / tot, 100 * s / max(ts, 1), f, ln, src[:100]))
