timeout 900 python -m pytest tests/test_gpu_audit.py tests/test_gpu_checkpoint.py tests/test_gpu_parity.py -x -q 2>&1 | tail -3
for w in dgauss rosen2d; do
timeout 200 python bench.py --steps 400 --warmup 10 --no-cpu --no-e2e --workload $w 2>/dev/null | python -c "import json,sys; j=json.loads(sys.stdin.read()); print('$w', '%.4f ms  %.4g'%(j['ms_per_step'], j['value']), j['roofline'].get('remote_iterations_mean'), j['accept_rate'], j['posterior_mean'])"
done
