#!/bin/bash
# round 2, final single-GPU measurement pass: GPU suite, bench lines of every workload, ncu metric lists (fp64 work, pipe utilisation,
# DRAM bytes), ncu --set full captures summarised on the box (reports stay there: 64 MiB limit on gpurun_out), launch list.
# Every command that runs under ncu has run plain first (same arguments) in this script.
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/j_smi.txt 2>&1
( time timeout 1800 python -m pytest tests -m gpu -q --timeout 900 ) > $O/j_tests.log 2>&1
echo "pytest rc=$?" >> $O/j_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/j_smoke.log 2>&1
( time timeout 900 python bench.py ) > $O/j_bench_dgauss.json 2> $O/j_bench_dgauss.err
( time timeout 600 python bench.py --impl reference ) > $O/j_bench_reference_arm.json 2> $O/j_bench_reference_arm.err
timeout 600 python bench.py --steps 20 --no-cpu --no-e2e --no-modes > $O/j_bench_dgauss_20steps.json 2>> $O/j_err.log
timeout 600 python bench.py --remote-mode summix --no-cpu > $O/j_bench_dgauss_summix.json 2>> $O/j_err.log
for w in rosen2d rosen16 gmix64; do
  timeout 900 python bench.py --workload $w --steps 200 > $O/j_bench_$w.json 2>> $O/j_err.log
done
timeout 600 python bench.py --workload gmix64 --remote-mode summix --pool 256 --steps 200 --no-cpu > $O/j_bench_gmix64_summix256.json 2>> $O/j_err.log
timeout 600 python bench.py --workload rosen16 --remote-mode summix --pool 256 --steps 200 --no-cpu --no-modes > $O/j_bench_rosen16_summix256.json 2>> $O/j_err.log
timeout 600 python bench.py --workload rosen2d --remote-mode summix --pool 256 --steps 200 --no-cpu --no-modes > $O/j_bench_rosen2d_summix256.json 2>> $O/j_err.log
MCGPU_NO_COOP=1 timeout 600 python bench.py --workload gmix64 --remote-mode summix --pool 256 --steps 100 --no-cpu --no-e2e --no-modes > $O/j_bench_gmix64_summix256_widekernel.json 2>> $O/j_err.log
# ---- ncu metric lists
M="smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum"
B="python bench.py --no-cpu --no-e2e --no-modes --advance 300 --steps 60"
G="python bench.py --workload gmix64 --no-cpu --no-e2e --no-modes --advance 60 --steps 40"
R="python bench.py --workload rosen16 --no-cpu --no-e2e --no-modes --advance 60 --steps 40"
ops() { tag=$1; kn=$2; skip=$3; cnt=$4; shift 4
  "$@" > $O/j_plain_$tag.json 2>> $O/j_err.log &&
  ncu --metrics $M --clock-control none -k regex:$kn -s $skip -c $cnt --csv --log-file $O/j_ops_$tag.csv "$@" > $O/j_ncu_ops_$tag.log 2>&1; }
ops ref16 mh_steps_kernel 264 100 $B --remote-mode reference --pool 16
ops sum16 mh_steps_kernel 264 100 $B --remote-mode summix --pool 16
ops sum256 mh_steps_kernel 264 100 $B --remote-mode summix --pool 256
ops local mh_steps_kernel 264 100 $B --pl 1.0
ops g64local mh_coop 30 40 $G --pl 1.0
ops g64sum256 mh_coop 30 60 $G --remote-mode summix --pool 256
ops g64ref16 "mh_coop|mh_wide" 30 60 $G
ops r16local mh_wide 30 40 $R --pl 1.0
ops r16ref16 mh_wide 30 40 $R
ops r16sum16 mh_wide 30 40 $R --remote-mode summix
# ---- ncu --set full, summarised here
full() { tag=$1; kn=$2; skip=$3; div=$4; shift 4
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$kn" -s $skip -c 3 -o /tmp/prof_$tag -f "$@" > $O/j_ncu_full_$tag.log 2>&1
  python tools/summarize_profile.py full /tmp/prof_$tag.ncu-rep $div $div $div > $O/r02_full_$tag.txt 2>> $O/j_err.log
  python tools/profile_lines.py /tmp/prof_$tag.ncu-rep 0 $div 70 > $O/r02_lines_$tag.txt 2>> $O/j_err.log; }
full local mh_steps_kernel 300 327680 $B --pl 1.0
full ref16 "mh_steps_kernel<3, 2, 0, 1>" 200 327680 $B --remote-mode reference --pool 16
full sum16 "mh_steps_kernel<3, 2, 0, 4>" 200 327680 $B --remote-mode summix --pool 16
full g64local "mh_coop_kernel<64, 4, 2>" 30 32768 $G --pl 1.0
full g64sum256 "mh_coop_kernel<64, 1, 5>" 3 131072 $G --remote-mode summix --pool 256
# ---- launch list of the default bench command
python bench.py --no-cpu --steps 100 > $O/j_plain_launches.json 2>> $O/j_err.log &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/j_launches.csv python bench.py --no-cpu --steps 100 > $O/j_ncu_launches.log 2>&1
python tools/summarize_profile.py launches $O/j_launches.csv > $O/r02_launches.txt 2>> $O/j_err.log
rm -f $O/j_launches.csv
find $O -size +12M -delete
du -sm $O
grep -E "passed|failed" $O/j_tests.log | tail -2; grep -E "^FAILED|^ERROR" $O/j_tests.log | head -20; cat $O/j_smoke.log | tail -2
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/j_bench_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print("%-40s %.4g  %.4f ms  iters %.1f  e2e %s  frac %s" % (f.split("/")[-1][8:-5], d["value"], d.get("ms_per_step", 0), d.get("remote_iterations_mean", 0),
              d.get("e2e") and "%.4g" % d["e2e"]["value"], d.get("roofline") and "%.3f" % d["roofline"]["frac"]))
        for k, v in (d.get("modes") or {}).items(): print("      mode %-14s %.4g  %.4f ms" % (k, v["value"], v["ms_per_step"]))
    except Exception as ex:
        print(f, "ERR", ex)
PY
tail -5 $O/j_err.log
