#!/bin/bash
# GPU-box experiment: launch-bound / CTA-size sweep of the step kernels (writes gpurun_out/tune.log)
set -u
out=gpurun_out/tune.log; : > $out
for minb in 1 5 6 8; do
  MCGPU_NVCC_FLAGS="-DMCGPU_MINB=$minb" python -m mcpar_b200.build --force > /dev/null 2>&1 || { echo "build failed minb=$minb" >> $out; continue; }
  for blk in 64 128; do
    for wl in dgauss rosen2d; do
      r=$(MCGPU_BLOCK=$blk python bench.py --steps 200 --warmup 5 --no-cpu --no-e2e --workload $wl 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.4g %.4f'%(d['value'], d['ms_per_step']))")
      echo "minb=$minb block=$blk $wl value ms_per_step: $r" >> $out
    done
  done
done
python -m mcpar_b200.build --force > /dev/null 2>&1
cat $out
