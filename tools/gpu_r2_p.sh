#!/bin/bash
# round 2, GPU call P (1 GPU): ncu --set full of the wide kernel on config 3 (Rosenbrock d = 16), local and reference-mode remote launches
mkdir -p gpurun_out
O=gpurun_out
R="python bench.py --workload rosen16 --no-cpu --no-e2e --no-modes --advance 60 --steps 40 --no-place"
full() { tag=$1; kn=$2; skip=$3; div=$4; shift 4
  "$@" > $O/p_plain_$tag.json 2>> $O/p_err.log &&
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$kn" -s $skip -c 3 -o /tmp/prof_$tag -f "$@" > $O/p_ncu_full_$tag.log 2>&1
  python tools/summarize_profile.py full /tmp/prof_$tag.ncu-rep $div $div $div > $O/r02_full_$tag.txt 2>> $O/p_err.log
  python tools/profile_lines.py /tmp/prof_$tag.ncu-rep 0 $div 60 > $O/r02_lines_$tag.txt 2>> $O/p_err.log; }
full r16local "mh_wide_kernel<[^>]*2>" 30 2621440 $R --pl 1.0
full r16ref16 "mh_wide_kernel<[^>]*3>" 20 262144 $R
sed -n 3,26p $O/r02_full_r16local.txt | cut -c1-250; sed -n 3,26p $O/r02_full_r16ref16.txt | cut -c1-250; tail -3 $O/p_err.log
