#!/bin/bash
# round 2, multi-GPU call (NG = $1 GPUs of one box): sharded == single-engine checks at 2/4/NG ranks, BASELINE config 5
# (chain-count x GPU-count sweep), config 4 (GaussMix d=64 K=64, 2^20 chains per GPU, exchange every sweep, pool M=256)
# and the default bench line (config 2) on NG GPUs.  Every launch is one torchrun; nothing here runs under ncu.
NG=${1:-8}
QUICK=${2:-0}
mkdir -p gpurun_out
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > $O/e${NG}_smi.txt 2>&1
port=29500
for w in 2 4 8; do
  [ $w -le $NG ] || continue
  port=$((port + 1))
  ( time timeout 600 $TR --nproc-per-node $w --master-port $port tests/dist_check.py ) > $O/e${NG}_dist_check_${w}gpu.log 2>&1
  echo "rc=$?" >> $O/e${NG}_dist_check_${w}gpu.log
done
if [ "$QUICK" = "1" ]; then SW="--min-exp 14 --max-exp 20 --ref-max-exp 16"; ST="--steps 50"; else SW="--ref-max-exp 20"; ST="--steps 200"; fi
port=$((port + 1))
( time timeout 900 $TR --nproc-per-node $NG --master-port $port tools/sweep_c5.py $SW ) > $O/e${NG}_sweep.log 2>&1
echo "rc=$?" >> $O/e${NG}_sweep.log
port=$((port + 1))
( time timeout 600 $TR --nproc-per-node $NG --master-port $port bench.py --gpus $NG --workload gmix64 --pool 256 --remote-mode summix $ST ) > $O/e${NG}_c4_summix256.json 2> $O/e${NG}_c4_summix256.err
echo "rc=$?" >> $O/e${NG}_c4_summix256.err
port=$((port + 1))
( time timeout 400 $TR --nproc-per-node $NG --master-port $port bench.py --gpus $NG --workload gmix64 --pool 256 --steps 60 --advance 60 --no-e2e --no-modes --no-check ) > $O/e${NG}_c4_reference256.json 2> $O/e${NG}_c4_reference256.err
echo "rc=$?" >> $O/e${NG}_c4_reference256.err
port=$((port + 1))
( time timeout 600 $TR --nproc-per-node $NG --master-port $port bench.py --gpus $NG $ST ) > $O/e${NG}_c2_dgauss.json 2> $O/e${NG}_c2_dgauss.err
echo "rc=$?" >> $O/e${NG}_c2_dgauss.err
if [ "$QUICK" = "2" ]; then
  port=$((port + 1))
  ( time timeout 600 $TR --nproc-per-node $NG --master-port $port bench.py --gpus $NG --remote-mode summix --no-modes $ST ) > $O/e${NG}_c2_dgauss_summix.json 2> $O/e${NG}_c2_dgauss_summix.err
  port=$((port + 1))
  ( time timeout 600 $TR --nproc-per-node $NG --master-port $port bench.py --gpus $NG --workload rosen16 --no-modes --steps 200 ) > $O/e${NG}_c3_rosen16.json 2> $O/e${NG}_c3_rosen16.err
fi
grep -h "dist_check\|rc=" $O/e${NG}_dist_check_*gpu.log | sort | uniq -c | sort -rn | head -30
tail -5 $O/e${NG}_sweep.log
cat $O/sweep_c5_${NG}gpu.md 2>/dev/null
for f in $O/e${NG}_c4_summix256 $O/e${NG}_c4_reference256 $O/e${NG}_c2_dgauss $O/e${NG}_c2_dgauss_summix $O/e${NG}_c3_rosen16; do
  [ -f $f.json ] || continue
  python - $f <<'PY'
import json, sys
f = sys.argv[1]
try:
    d = json.loads([l for l in open(f + ".json") if l.startswith("{")][-1])
    print(f.split("/")[-1], "value %.4g  ms %.4f  per-rank %s  wait %.4f ms  iters %.1f  e2e %s  check %s  mean %s" % (
        d["value"], d["ms_per_step"], ["%.4f" % x for x in d["per_rank_ms_per_step"]], d["exchange_wait_ms_per_step"], d["remote_iterations_mean"],
        d["e2e"] and "%.4g" % d["e2e"]["value"], d["sharded_equals_single"], ["%.3f" % x for x in d["posterior_mean"][:2]]))
except Exception as ex:
    print(f, "ERR", ex); print(open(f + ".err").read()[-1500:])
PY
done
