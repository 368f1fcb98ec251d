mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/f_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/f_gpu_tests.log
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/f_bench.json
timeout 600 python bench.py --impl reference > gpurun_out/f_ref.json 2> gpurun_out/f_ref.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/f_ref.json
for w in rosen2d rosen16 gmix64; do timeout 300 python bench.py --steps 200 --warmup 5 --no-cpu --no-e2e --workload $w > gpurun_out/f_$w.json 2>/dev/null; echo "$w $(cut -c1-140 gpurun_out/f_$w.json)"; done
timeout 200 python bench.py --steps 400 --warmup 10 --no-cpu --no-e2e --pl 1.0 > gpurun_out/f_local.json 2>/dev/null; cut -c1-140 gpurun_out/f_local.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/f_launches.csv python bench.py --steps 300 --warmup 3 --no-cpu --no-e2e > gpurun_out/f_ncu1.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mh_steps -s 262 -c 6 -o gpurun_out/prof_r01d -f python bench.py --steps 300 --warmup 3 --no-cpu --no-e2e > gpurun_out/f_ncu2.log 2>&1; echo "ncu full rc=$?"; ls -la gpurun_out/prof_r01d.ncu-rep
