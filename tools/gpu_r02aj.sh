#!/bin/bash
O=gpurun_out/r02aj
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc $?" >> $O/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1
timeout 600 python bench.py > $O/bench_c3.json 2> $O/bench_c3.err
tail -3 $O/pytest_gpu.log; tail -1 $O/smoke.log; cut -c1-200 $O/bench_c3.json; tail -2 $O/bench_c3.err
