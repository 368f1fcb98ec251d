#!/usr/bin/env bash
# Reproduces the measurements committed under profiles/ (run on a B200 box, e.g. under gpurun):
#   tools/profile_round.sh r02        # writes gpurun_out/r02_* ; summarise here with tools/summarize_profile.py
# Every number is taken from a run WITHOUT ncu; the ncu passes follow only after that command exited 0.
set -u
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
python bench.py > $OUT/${TAG}_bench_dgauss_1gpu.json 2> $OUT/${TAG}_bench.err || exit 1
python bench.py --impl reference > $OUT/${TAG}_bench_reference_arm.json 2>> $OUT/${TAG}_bench.err || exit 1
for w in rosen2d rosen16 gmix64; do
  python bench.py --steps 200 --warmup 5 --no-cpu --no-e2e --workload $w > $OUT/${TAG}_bench_$w.json 2>/dev/null
done
python bench.py --steps 400 --warmup 10 --no-cpu --no-e2e --pl 1.0 > $OUT/${TAG}_bench_local_only.json 2>/dev/null
# launch list (cold-cache, serialised: compare shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 300 --warmup 3 --no-cpu --no-e2e > $OUT/${TAG}_ncu1.log 2>&1
# full capture of late windows (10 burn-in launches + 3 warm-up + 249 windows are skipped: windows 252-257)
ncu --set full --clock-control none --import-source on -k regex:mh_steps -s 262 -c 6 -o $OUT/prof_${TAG} -f \
    python bench.py --steps 300 --warmup 3 --no-cpu --no-e2e > $OUT/${TAG}_ncu2.log 2>&1
echo "then, in the dev container:"
echo "  python tools/summarize_profile.py launches $OUT/${TAG}_launches.csv > profiles/${TAG}_launches.txt"
echo "  python tools/summarize_profile.py full $OUT/prof_${TAG}.ncu-rep 327680 327680 327680 327680 327680 327680 > profiles/${TAG}_full.txt"
echo "  python tools/profile_lines.py $OUT/prof_${TAG}.ncu-rep 0 32768 60 > profiles/${TAG}_lines_window252.txt"
