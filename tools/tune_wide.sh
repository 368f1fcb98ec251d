#!/bin/bash
# GPU-box experiment: chains-per-group / launch-bound sweep of the wide kernel at d = 64
out=gpurun_out/tune_wide.log; : > $out
for nch in 1 2 4; do for minb in 2 3 4; do
  MCGPU_NVCC_FLAGS="-DMCGPU_WIDE_NCH64=$nch -DMCGPU_WIDE_MINB=$minb" python -m mcpar_b200.build --force > /dev/null 2>&1 || { echo "build failed" >> $out; continue; }
  r=$(python bench.py --steps 60 --warmup 5 --no-cpu --no-e2e --workload gmix64 --pl 1.0 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.4g %.4f'%(d['value'], d['ms_per_step']))")
  echo "nch=$nch minb=$minb gmix64 pl=1: $r" >> $out
done; done
python -m mcpar_b200.build --force > /dev/null 2>&1
cat $out
