#!/bin/bash
O=gpurun_out/r02t
mkdir -p $O
timeout 600 python tests/bench_ncc.py 100 200 500 1000 2000 > $O/bench_ncc.txt 2>&1
timeout 600 ncu --set full --clock-control none -k regex:k_search_ncc --launch-skip 2 --launch-count 2 -f -o $O/prof_ncc python tests/bench_ncc.py 2000 > $O/ncu_ncc.log 2>&1
cat $O/bench_ncc.txt; tail -3 $O/ncu_ncc.log
