#!/bin/bash
# round 2, call AM (4 GPUs): the driver's scaling command at N = 4 and N = 2, final code
O=gpurun_out/r02am
mkdir -p $O
for n in 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n bench.py --gpus $n --steps 20 --warmup 5 > $O/bench_${n}gpu.json 2> $O/bench_${n}gpu.err
done
for n in 4 2; do cut -c1-200 $O/bench_${n}gpu.json; tail -1 $O/bench_${n}gpu.err; done
