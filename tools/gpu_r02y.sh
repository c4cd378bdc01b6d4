#!/bin/bash
O=gpurun_out/r02y
mkdir -p $O
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_ncc.py tests/test_gpu_map.py tests/test_gpu_edges.py -m gpu -x -q > $O/sanitizer_memcheck.log 2>&1; echo "memcheck rc $?" >> $O/sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lanes or batched or multi_block or numeric or downdate_kernel" > $O/sanitizer_memcheck2.log 2>&1; echo "memcheck rc $?" >> $O/sanitizer_memcheck2.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_ncc.py -m gpu -x -q -k "capture or step_with" > $O/sanitizer_racecheck.log 2>&1; echo "racecheck rc $?" >> $O/sanitizer_racecheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "multi_block or downdate_kernel and 313" > $O/sanitizer_racecheck2.log 2>&1; echo "racecheck rc $?" >> $O/sanitizer_racecheck2.log
for f in $O/sanitizer_*.log; do echo $f; tail -4 $f; done
