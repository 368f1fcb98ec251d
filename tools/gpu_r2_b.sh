#!/bin/bash
# round 2, GPU call B (1 GPU): full GPU test suite, default bench, ncu captures, bias curve, C3/C4 single-GPU, C5 R=1 column
mkdir -p gpurun_out
O=gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -q --timeout 900 ) > $O/b_tests.log 2>&1
echo "pytest rc=$?" >> $O/b_tests.log
( time timeout 900 python bench.py ) > $O/b_bench.json 2> $O/b_bench.err
echo "bench rc=$?" >> $O/b_bench.err
timeout 600 python tools/bias_curve.py > $O/b_bias.log 2>&1
for w in rosen2d rosen16 gmix64; do
  timeout 600 python bench.py --workload $w --steps 200 --no-cpu > $O/b_bench_$w.json 2>> $O/b_bench.err
done
timeout 600 python bench.py --workload gmix64 --remote-mode summix --pool 256 --steps 100 --no-cpu --no-modes --no-e2e > $O/b_bench_gmix64_summix256.json 2>> $O/b_bench.err
timeout 600 python bench.py --workload rosen16 --remote-mode summix --pool 256 --steps 100 --no-cpu --no-modes --no-e2e > $O/b_bench_rosen16_summix256.json 2>> $O/b_bench.err
timeout 900 python tools/sweep_c5.py --ref-max-exp 20 > $O/b_sweep.log 2>&1
# ---- ncu (every profiled command line ran above or runs plain first).  The reports are summarised HERE and only
# the text comes home: gpurun_out/ may not exceed 64 MiB (four reports with imported source were 140 MB).
B="python bench.py --no-cpu --no-e2e --no-modes --advance 300 --steps 60"
M="smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum"
for cfg in "ref16:--remote-mode reference --pool 16" "sum16:--remote-mode summix --pool 16" "sum256:--remote-mode summix --pool 256" "local:--pl 1.0"; do
  tag=${cfg%%:*}; fl=${cfg#*:}
  $B $fl > $O/b_plain_$tag.json 2>> $O/b_bench.err &&
  ncu --metrics $M --clock-control none -k regex:mh_steps_kernel -s 264 -c 100 --csv --log-file $O/b_ops_$tag.csv $B $fl > $O/b_ncu_ops_$tag.log 2>&1
  ncu --set full --clock-control none -k regex:mh_steps_kernel -s 350 -c 4 -o /tmp/prof_r02_$tag -f $B $fl > $O/b_ncu_full_$tag.log 2>&1
  python tools/summarize_profile.py full /tmp/prof_r02_$tag.ncu-rep 327680 327680 327680 327680 > $O/r02_full_$tag.txt 2>> $O/b_bench.err
  for k in 0 1 2 3; do python tools/profile_lines.py /tmp/prof_r02_$tag.ncu-rep $k 327680 70 > $O/r02_lines_${tag}_$k.txt 2>> $O/b_bench.err; done
  ncu -i /tmp/prof_r02_$tag.ncu-rep --page details --csv > $O/r02_details_$tag.csv 2>> $O/b_bench.err
  ls -la /tmp/prof_r02_$tag.ncu-rep >> $O/b_bench.err
  [ $(stat -c %s /tmp/prof_r02_$tag.ncu-rep) -lt 9000000 ] && cp /tmp/prof_r02_$tag.ncu-rep $O/
done
$B > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $O/b_launches.csv $B > $O/b_ncu_launches.log 2>&1
find $O -size +12M -delete
du -sm $O
ls -la $O | grep -E " b_| r02_|prof_r02" | awk '{print $5, $9}'
grep -E "passed|failed" $O/b_tests.log | tail -3
grep -E "^FAILED|^ERROR" $O/b_tests.log | head -20
