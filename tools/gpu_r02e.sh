#!/bin/bash
# round 2, call E (2 GPUs): multi-GPU path of the new bench.py, chain phase clocks, vendor bar with the mirror pass
O=gpurun_out/r02e
mkdir -p $O
timeout 200 python tools/dbg_chain.py > $O/dbg_chain.txt 2>&1
timeout 120 tools/vendor_bar 3013 640 1 3013 72 1 3013 1000 1 1213 290 32 > $O/vendor_bar.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_c3_2gpu.json 2> $O/bench_c3_2gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 5 --workload c2 > $O/bench_ref_c2_2gpu.json 2> $O/bench_ref_c2_2gpu.err
cat $O/dbg_chain.txt; cat $O/vendor_bar.txt; cut -c1-2500 $O/bench_c3_2gpu.json; tail -5 $O/bench_c3_2gpu.err; cut -c1-400 $O/bench_ref_c2_2gpu.json
