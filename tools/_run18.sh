for b in 64 128 32; do
for pl in 0.9 1.0; do
MCGPU_BLOCK=$b timeout 200 python bench.py --steps 400 --warmup 10 --no-cpu --no-e2e --pl $pl 2>/dev/null | python -c "import json,sys; j=json.loads(sys.stdin.read()); print('block $b pl $pl', '%.4f ms  %.4g'%(j['ms_per_step'], j['value']))"
done; done
