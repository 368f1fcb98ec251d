#!/bin/bash
# round 2, GPU call A (1 GPU): GPU test suite, smoke, default bench, launch-bound A/B of the local kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/a_smi.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 ) > gpurun_out/a_tests.log 2>&1
echo "pytest rc=$?" >> gpurun_out/a_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/a_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/a_smoke.log
( time timeout 900 python bench.py ) > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err
echo "bench rc=$?" >> gpurun_out/a_bench.err
for v in mb10 mb9 mb8 mb7; do
  MCGPU_LIB=$PWD/mcpar_b200/variants/libmcgpu_$v.so timeout 300 python bench.py --pl 1.0 --no-cpu --no-e2e --no-modes --steps 300 --advance 100 > gpurun_out/a_local_$v.json 2>> gpurun_out/a_bench.err
  MCGPU_LIB=$PWD/mcpar_b200/variants/libmcgpu_$v.so timeout 300 python bench.py --no-cpu --no-e2e --no-modes --steps 300 --advance 1000 > gpurun_out/a_win_$v.json 2>> gpurun_out/a_bench.err
done
tail -3 gpurun_out/a_tests.log
