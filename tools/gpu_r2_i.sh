#!/bin/bash
# round 2, GPU call I (1 GPU): trimmed cooperative kernel -- parity subset, bench A/B, ncu of the remote (sum-mixture M = 256) kernel with stall reasons
mkdir -p gpurun_out
O=gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 300 -k "gaussmix or gmix" ) > $O/i_tests_gmix.log 2>&1
echo "pytest rc=$?" >> $O/i_tests_gmix.log
grep -E 'passed|failed' $O/i_tests_gmix.log | tail -2; grep -E '^FAILED|^ERROR' $O/i_tests_gmix.log | head
S="--workload gmix64 --steps 100 --advance 100 --no-cpu --no-e2e --no-modes"
run() { tag=$1; shift; timeout 300 python bench.py $S "$@" > $O/i_$tag.json 2>> $O/i_err.log || echo "FAILED $tag" >> $O/i_err.log; }
run coop_local --pl 1.0; run coop_sum256 --remote-mode summix --pool 256; run coop_sum16 --remote-mode summix --pool 16; run coop_ref16
run coop_allremote256 --remote-mode summix --pool 256 --pl 0.0; run coop_allremote16 --remote-mode summix --pool 16 --pl 0.0
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/i_*.json")):
    try:
        d = json.load(open(f)); print("%-28s %.4g  %.4f ms  fallback %.2e  accept %.3f  remote %.3f" % (f.split("/")[-1][2:-5], d["value"], d["ms_per_step"], d["exact_fallback_rate"], d["accept_rate"], d["remote_fraction"]))
    except Exception as ex:
        print(f, "ERR", ex)
PY
tail -20 $O/i_err.log
B="python bench.py --workload gmix64 --no-cpu --no-e2e --no-modes --advance 20 --steps 20"
for cfg in "g64sum256:--remote-mode summix --pool 256 --pl 0.0:16" "g64local:--pl 1.0:30"; do
  tag=${cfg%%:*}; rest=${cfg#*:}; fl=${rest%%:*}; skip=${rest#*:}
  $B $fl > $O/i_plain_$tag.json 2>> $O/i_err.log &&
  ncu --set full --clock-control none --import-source on -k regex:mh_coop -s $skip -c 2 -o /tmp/prof_$tag -f $B $fl > $O/i_ncu_$tag.log 2>&1
  python tools/summarize_profile.py full /tmp/prof_$tag.ncu-rep 65536 65536 > $O/r02_full_$tag.txt 2>> $O/i_err.log
  python tools/profile_lines.py /tmp/prof_$tag.ncu-rep 0 65536 120 > $O/r02_lines_$tag.txt 2>> $O/i_err.log
  ncu -i /tmp/prof_$tag.ncu-rep --page raw --csv --launch-count 1 | python -c "
import csv, sys
rows = list(csv.reader(sys.stdin))
h, v = rows[0], rows[2]
for a, b in zip(h, v):
    if 'stall' in a or 'issue' in a or 'warp_latency' in a: print(a, b)
" > $O/r02_stalls_$tag.txt 2>> $O/i_err.log
done
head -26 $O/r02_full_g64sum256.txt; cat $O/r02_stalls_g64sum256.txt | sort -t' ' -k2 -rn | head -30
