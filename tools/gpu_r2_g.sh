#!/bin/bash
# round 2, GPU call G (1 GPU): ncu --set full of the cooperative d = 64 kernels (local and sum-mixture remote), summarised on the box
mkdir -p gpurun_out
O=gpurun_out
B="python bench.py --workload gmix64 --no-cpu --no-e2e --no-modes --advance 20 --steps 20"
for cfg in "g64local:--pl 1.0:mh_coop:30" "g64sum256:--remote-mode summix --pool 256 --pl 0.5:mh_coop_kernel<64, 2, 5>:4"; do
  tag=${cfg%%:*}; rest=${cfg#*:}; fl=${rest%%:*}; rest=${rest#*:}; kn=${rest%%:*}; skip=${rest#*:}
  $B $fl > $O/g_plain_$tag.json 2>> $O/g_err.log &&
  ncu --set full --clock-control none --import-source on -k "regex:$kn" -s $skip -c 2 -o /tmp/prof_$tag -f $B $fl > $O/g_ncu_$tag.log 2>&1
  python tools/summarize_profile.py full /tmp/prof_$tag.ncu-rep 65536 65536 > $O/r02_full_$tag.txt 2>> $O/g_err.log
  python tools/profile_lines.py /tmp/prof_$tag.ncu-rep 0 65536 90 > $O/r02_lines_$tag.txt 2>> $O/g_err.log
  ncu -i /tmp/prof_$tag.ncu-rep --page details --csv --launch-count 1 > $O/r02_details_$tag.csv 2>> $O/g_err.log
  ls -la /tmp/prof_$tag.ncu-rep >> $O/g_err.log
done
head -40 $O/r02_full_g64local.txt
head -50 $O/r02_lines_g64local.txt
grep -i "stall\|No Eligible\|Eligible Warps\|Issued Warp\|Active Warps" $O/r02_details_g64local.csv | cut -c1-260 | head -40
