#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the text summaries committed under profiles/.

  python tools/summarize_profile.py launches gpurun_out/launches.csv > profiles/r01_launches.txt
  python tools/summarize_profile.py full gpurun_out/prof.ncu-rep [warp-steps per launch ...] > profiles/r01_full.txt
"""
import collections
import csv
import re
import subprocess
import sys

RAW = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
       "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
       "smsp__issue_active.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
       "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
       "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
       "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
       "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_elapsed.max"]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    order = []
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(row["Metric Unit"], 1e-3)
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        agg.setdefault(name, []).append(v)
        order.append((name, v))
    tot = sum(sum(v) for v in agg.values())
    print("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)")
    print("%-70s %6s %12s %10s %7s" % ("kernel", "n", "total us", "avg us", "share"))
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print("%-70s %6d %12.1f %10.1f %6.1f%%" % (k[:70], len(v), sum(v), sum(v) / len(v), 100 * sum(v) / tot))


def full(path, warpsteps):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    iname = hdr.index("Kernel Name")
    print("# ncu --set full --clock-control none (raw page); one block per captured launch")
    for k, row in enumerate(rows[2:]):
        print("\n## launch %d: %s" % (k, row[iname][:100]))
        for m in RAW:
            if m in hdr:
                i = hdr.index(m)
                print("  %-70s %s %s" % (m, row[i], units[i]))
        # warp stall reasons (cycles a warp spent stalled per instruction it issued), largest first
        st = []
        for i, m in enumerate(hdr):
            if "issue_stalled" in m and m.endswith("_per_warp_active.pct"):
                try:
                    st.append((float(row[i]), m.replace("smsp__average_warps_issue_stalled_", "").replace("_per_warp_active.pct", "")))
                except ValueError:
                    pass
        if not st:
            for i, m in enumerate(hdr):
                if "issue_stalled" in m and m.endswith(".ratio"):
                    try:
                        st.append((float(row[i]), m.replace("smsp__average_warp_latency_issue_stalled_", "").replace("smsp__average_warps_issue_stalled_", "").replace(".ratio", "")))
                    except ValueError:
                        pass
        if st:
            print("  warp stall reasons: " + ", ".join("%s %.2f" % (n, v) for v, n in sorted(st, reverse=True)[:8]))
        src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--launch-skip", str(k), "--launch-count", "1"],
                             capture_output=True, text=True).stdout
        srows = list(csv.reader(src.splitlines()))
        if len(srows) < 3:
            continue
        sh = srows[1]
        iS, iE, iSm = sh.index("Source"), sh.index("Instructions Executed"), sh.index("# Samples")
        tot, samp = collections.Counter(), collections.Counter()
        for r in srows[2:]:
            if len(r) <= iE or r[iE] == "Instructions Executed":
                continue
            ins = re.sub(r"^@!?U?P\w+\s+", "", r[iS].strip())
            if not ins:
                continue
            op = ins.split()[0].split(".")[0]
            try:
                tot[op] += float(r[iE] or 0) / 2; samp[op] += float(r[iSm] or 0) / 2   # page lists each instruction twice
            except ValueError:
                pass
        n = sum(tot.values())
        ws = warpsteps[k] if k < len(warpsteps) else None
        print("  SASS mix (warp instructions%s):" % (", per warp-step" if ws else ""))
        for op, v in tot.most_common(16):
            print("    %-8s %14.0f %6.1f%%%s   stall samples %6.0f" % (op, v, 100 * v / n, ("  %7.1f" % (v / ws)) if ws else "", samp[op]))
        if ws:
            print("    total per warp-step: %.1f" % (n / ws))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        full(sys.argv[2], [float(x) for x in sys.argv[3:]])
