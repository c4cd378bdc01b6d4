#!/bin/bash
# round 2, call R (8 GPUs): the driver's scaling commands at N = 8, 4, 2 and 1 on one box (C3 replicas + the C4 sharded leg in one line)
O=gpurun_out/r02r
mkdir -p $O
for n in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 20 --warmup 5 > $O/bench_${n}gpu.json 2> $O/bench_${n}gpu.err
done
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_1gpu.json 2> $O/bench_1gpu.err
for n in 8 4 2 1; do cut -c1-300 $O/bench_${n}gpu.json; tail -2 $O/bench_${n}gpu.err; done
