#!/bin/bash
O=gpurun_out/r02aa
mkdir -p $O
timeout 600 python -m pytest tests/test_ncc.py tests/test_frontend.py -m gpu -x -q > $O/pytest_ncc.log 2>&1; echo "pytest rc $?" >> $O/pytest_ncc.log
timeout 300 python tests/bench_ncc.py 100 500 2000 > $O/bench_ncc.txt 2>&1
EKFB_OPTS="16=0" timeout 300 python tests/bench_ncc.py 100 500 2000 > $O/bench_ncc_bulk.txt 2>&1
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_ncc.py -m gpu -x -q -k "tensor_map or equals_oracle" > $O/sanitizer_ncc.log 2>&1; echo "memcheck rc $?" >> $O/sanitizer_ncc.log
tail -5 $O/pytest_ncc.log; head -3 $O/bench_ncc.txt | cut -c1-200; head -3 $O/bench_ncc_bulk.txt | cut -c1-200; tail -3 $O/sanitizer_ncc.log
