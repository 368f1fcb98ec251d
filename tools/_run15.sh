MCGPU_LIB=$PWD/ab/libmcgpu_v26.so timeout 900 python -m pytest tests/test_gpu_audit.py tests/test_gpu_parity.py -x -q 2>&1 | tail -3
for v in v23 v25 v26; do
for w in dgauss rosen2d; do
MCGPU_LIB=$PWD/ab/libmcgpu_$v.so timeout 200 python bench.py --steps 400 --warmup 10 --no-cpu --no-e2e --workload $w 2>/dev/null | python -c "import json,sys; j=json.loads(sys.stdin.read()); print('$v $w', '%.4f ms  %.4g'%(j['ms_per_step'], j['value']), j['roofline'].get('remote_iterations_mean'), j['accept_rate'], j['posterior_mean'])"
done; done
