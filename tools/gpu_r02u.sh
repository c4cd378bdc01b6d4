#!/bin/bash
O=gpurun_out/r02u
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_parity_bench_sizes.py -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc $?" >> $O/pytest_gpu.log
timeout 200 python tools/quick_time.py 640 480 500 1 80 > $O/quick_c3.txt 2>&1
timeout 200 python tools/dbg_chain.py > $O/dbg_chain.txt 2>&1
timeout 200 python tools/quick_time.py 640 480 200 1 60 > $O/quick_n200.txt 2>&1
tail -3 $O/pytest_gpu.log; cat $O/dbg_chain.txt | head -3; for f in $O/quick_*.txt; do echo $f; tail -2 $f | cut -c1-420; done
