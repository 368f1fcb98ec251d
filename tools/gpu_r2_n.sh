#!/bin/bash
# round 2, GPU call N (1 GPU): validation of the last cooperative-kernel change (full pool: no totals pass) + config 4 lines + metric lists
mkdir -p gpurun_out
O=gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_p2p.py tests/test_gpu_checkpoint.py -m gpu -q --timeout 600 ) > $O/n_tests.log 2>&1
echo "pytest rc=$?" >> $O/n_tests.log
grep -E "passed|failed" $O/n_tests.log | tail -2; grep -E "^FAILED|^ERROR" $O/n_tests.log | head
timeout 900 python bench.py --workload gmix64 --steps 200 > $O/n_bench_gmix64.json 2>> $O/n_err.log
timeout 600 python bench.py --workload gmix64 --remote-mode summix --pool 256 --steps 200 --no-cpu > $O/n_bench_gmix64_summix256.json 2>> $O/n_err.log
M="smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum"
G="python bench.py --workload gmix64 --no-cpu --no-e2e --no-modes --advance 300 --steps 40 --no-place --remote-mode summix --pool 256"
$G > $O/n_plain_g64sum256.json 2>> $O/n_err.log &&
ncu --metrics $M --clock-control none -k regex:mh_coop -s 320 -c 60 --csv --log-file $O/n_ops_g64sum256.csv $G > $O/n_ncu_ops.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:mh_coop_kernel<[^>]*5, " -s 25 -c 3 -o /tmp/prof_g64sum256 -f $G > $O/n_ncu_full.log 2>&1
python tools/summarize_profile.py full /tmp/prof_g64sum256.ncu-rep 131072 131072 131072 > $O/r02_full_g64sum256.txt 2>> $O/n_err.log
python tools/profile_lines.py /tmp/prof_g64sum256.ncu-rep 0 131072 70 > $O/r02_lines_g64sum256.txt 2>> $O/n_err.log
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/n_bench_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print("%-32s %.4g  %.4f ms  remote %.3f  e2e %s  frac %s" % (f.split("/")[-1][8:-5], d["value"], d["ms_per_step"], d["remote_fraction"], d.get("e2e") and "%.4g" % d["e2e"]["value"], d.get("roofline") and "%.3f" % d["roofline"]["frac"]))
        for k, v in (d.get("modes") or {}).items(): print("      mode %-14s %.4g  %.4f ms" % (k, v["value"], v["ms_per_step"]))
    except Exception as ex:
        print(f, "ERR", ex)
PY
sed -n 3,10p $O/r02_full_g64sum256.txt | cut -c1-200; tail -3 $O/n_err.log
