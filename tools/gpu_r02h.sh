#!/bin/bash
# round 2, call H (8 GPUs): the driver's scaling commands at N = 8 and N = 4 (C3 replicas + the C4 sharded leg in one line)
O=gpurun_out/r02h
mkdir -p $O
for n in 8 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 20 --warmup 5 > $O/bench_${n}gpu.json 2> $O/bench_${n}gpu.err
done
for n in 8 4; do cut -c1-400 $O/bench_${n}gpu.json; tail -2 $O/bench_${n}gpu.err; done
