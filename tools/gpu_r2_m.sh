#!/bin/bash
# round 2, second multi-GPU call (NG GPUs of one box): config 4 with the cooperative kernel and config 2 with the final d = 2 kernels
NG=${1:-8}
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/m${NG}_smoke.log 2>&1; tail -1 $O/m${NG}_smoke.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $NG"
( time timeout 600 $TR --master-port 29611 bench.py --gpus $NG --workload gmix64 --pool 256 --remote-mode summix --steps 200 ) > $O/m${NG}_c4_summix256.json 2> $O/m${NG}_c4_summix256.err
echo "rc=$?" >> $O/m${NG}_c4_summix256.err
( time timeout 400 $TR --master-port 29612 bench.py --gpus $NG --workload gmix64 --pool 256 --steps 60 --advance 60 --no-e2e --no-modes --no-check ) > $O/m${NG}_c4_reference256.json 2> $O/m${NG}_c4_reference256.err
echo "rc=$?" >> $O/m${NG}_c4_reference256.err
( time timeout 600 $TR --master-port 29613 bench.py --gpus $NG --steps 200 ) > $O/m${NG}_c2_dgauss.json 2> $O/m${NG}_c2_dgauss.err
echo "rc=$?" >> $O/m${NG}_c2_dgauss.err
for f in $O/m${NG}_c4_summix256 $O/m${NG}_c4_reference256 $O/m${NG}_c2_dgauss; do
  python - $f <<'PY'
import json, sys
f = sys.argv[1]
try:
    d = json.loads([l for l in open(f + ".json") if l.startswith("{")][-1])
    print(f.split("/")[-1], "value %.4g  ms %.4f  per-rank %s  wait %.4f ms  iters %.1f  e2e %s  check %s  modes %s" % (
        d["value"], d["ms_per_step"], ["%.4f" % x for x in d["per_rank_ms_per_step"]], d["exchange_wait_ms_per_step"], d["remote_iterations_mean"],
        d["e2e"] and "%.4g" % d["e2e"]["value"], d["sharded_equals_single"], {k: "%.4g" % v["value"] for k, v in (d.get("modes") or {}).items()}))
except Exception as ex:
    print(f, "ERR", ex); print(open(f + ".err").read()[-1500:])
PY
done
