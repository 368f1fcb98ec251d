mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29515"
for x in p2p nccl; do
timeout 200 $TR --nproc-per-node 8 bench.py --gpus 8 --steps 400 --warmup 10 --no-cpu --no-e2e --exchange $x > gpurun_out/g8_$x.json 2> gpurun_out/g8_$x.err; echo "$x rc=$? $(cut -c1-170 gpurun_out/g8_$x.json)"
done
timeout 200 python bench.py --steps 400 --warmup 10 --no-cpu --no-e2e > gpurun_out/g1.json 2>/dev/null; cut -c1-170 gpurun_out/g1.json
