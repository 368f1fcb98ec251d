#!/bin/bash
# GPU-box experiment: chains-per-group / launch-bound / unroll sweep of the wide kernel at d = 64
out=gpurun_out/tune_wide2.log; : > $out
for cfg in "1 6 4" "1 8 4" "2 5 4" "2 6 4" "2 4 8" "4 3 8" "4 3 2" "2 5 8"; do
  set -- $cfg
  MCGPU_NVCC_FLAGS="-DMCGPU_WIDE_NCH64=$1 -DMCGPU_WIDE_MINB=$2 -DMCGPU_WIDE_UNROLL=$3" python -m mcpar_b200.build --force > /dev/null 2>&1 || { echo "build failed $cfg" >> $out; continue; }
  r=$(python bench.py --steps 60 --warmup 5 --no-cpu --no-e2e --workload gmix64 --pl 1.0 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.4g %.4f'%(d['value'], d['ms_per_step']))")
  echo "nch=$1 minb=$2 unroll=$3 gmix64 pl=1: $r" >> $out
done
python -m mcpar_b200.build --force > /dev/null 2>&1
cat $out
