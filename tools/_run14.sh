mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mh_steps -s 262 -c 6 -o gpurun_out/prof_r01c -f python bench.py --steps 300 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_c.log 2>&1; echo rc=$?; ls -la gpurun_out/prof_r01c.ncu-rep
