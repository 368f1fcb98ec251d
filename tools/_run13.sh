mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mh_steps -s 9 -c 6 -o gpurun_out/prof_r01b -f python bench.py --steps 30 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_b.log 2>&1; echo rc=$?; tail -3 gpurun_out/ncu_b.log; ls -la gpurun_out/prof_r01b.ncu-rep
