#!/bin/bash
# round 2, call F: chain with the deferred X publish, bench line with the per-update roofline split
O=gpurun_out/r02f
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q -k "not c4_shape and not s3_sequence_follows" > $O/pytest_gpu.log 2>&1; echo "pytest rc $?" >> $O/pytest_gpu.log
timeout 200 python tools/dbg_chain.py > $O/dbg_chain.txt 2>&1
timeout 200 python tools/quick_time.py 640 480 500 1 80 > $O/quick_c3.txt 2>&1
timeout 600 python bench.py --no-cpu-baseline > $O/bench_c3.json 2> $O/bench_c3.err
tail -3 $O/pytest_gpu.log; cat $O/dbg_chain.txt; tail -2 $O/quick_c3.txt | cut -c1-700; cut -c1-300 $O/bench_c3.json; tail -3 $O/bench_c3.err
