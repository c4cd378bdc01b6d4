#!/bin/bash
O=gpurun_out/r02j
mkdir -p $O
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py tests/test_parity_bench_sizes.py -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc $?" >> $O/pytest_gpu.log
timeout 200 python tools/downdate_sweep.py 500 72 640 1000 > $O/downdate_sweep.txt 2>&1
timeout 200 python tools/dbg_chain.py > $O/dbg_chain.txt 2>&1
timeout 200 python tools/quick_time.py 640 480 500 1 80 > $O/quick_c3.txt 2>&1
timeout 200 python tools/quick_time.py 320 240 50 1 80 > $O/quick_c2.txt 2>&1
timeout 200 python tools/quick_time.py 640 480 200 32 40 > $O/quick_c4_32.txt 2>&1
tail -3 $O/pytest_gpu.log; cat $O/dbg_chain.txt $O/downdate_sweep.txt; for f in $O/quick_*.txt; do echo $f; tail -2 $f | cut -c1-600; done
