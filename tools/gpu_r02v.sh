#!/bin/bash
O=gpurun_out/r02v
mkdir -p $O
timeout 200 python tools/quick_time.py 640 480 500 1 80 > $O/quick_c3.txt 2>&1
timeout 200 python tools/dbg_chain.py > $O/dbg_chain.txt 2>&1
timeout 200 python tools/quick_time.py 640 480 500 1 80 > $O/quick_c3_b.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "schain or factorisation or phase_by_phase or c3" > $O/pytest_gpu.log 2>&1; echo "pytest rc $?" >> $O/pytest_gpu.log
tail -3 $O/pytest_gpu.log; cat $O/dbg_chain.txt | head -1; for f in $O/quick_*.txt; do echo $f; tail -2 $f | cut -c1-420; done
