#!/bin/bash
# round 2, call X (8 GPUs): the driver's scaling command at N = 8 and N = 1 on one box, final code
O=gpurun_out/r02x
mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29538 bench.py --gpus 8 --steps 20 --warmup 5 > $O/bench_8gpu.json 2> $O/bench_8gpu.err
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_1gpu.json 2> $O/bench_1gpu.err
for n in 8 1; do cut -c1-200 $O/bench_${n}gpu.json; tail -2 $O/bench_${n}gpu.err; done
