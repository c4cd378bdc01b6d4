#!/bin/bash
O=gpurun_out/r02s
mkdir -p $O
timeout 600 python -m pytest tests/test_ncc.py tests/test_gpu_map.py tests/test_frontend.py -m gpu -x -q > $O/pytest_ncc.log 2>&1; echo "pytest rc $?" >> $O/pytest_ncc.log
timeout 300 python tests/bench_ncc.py 100 500 2000 > $O/bench_ncc.txt 2>&1
tail -25 $O/pytest_ncc.log; cat $O/bench_ncc.txt
