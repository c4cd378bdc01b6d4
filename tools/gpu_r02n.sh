#!/bin/bash
O=gpurun_out/r02n
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc $?" >> $O/pytest_gpu.log
timeout 200 python tools/dd_probe.py 500 72 640 > $O/dd_probe.txt 2>&1
timeout 200 python tools/downdate_sweep.py 500 72 640 1000 > $O/downdate_sweep.txt 2>&1
timeout 200 python tools/downdate_sweep.py 200 64 290 > $O/downdate_sweep_n200.txt 2>&1
timeout 200 python tools/quick_time.py 640 480 500 1 80 > $O/quick_c3.txt 2>&1
timeout 200 python tools/quick_time.py 640 480 200 32 40 > $O/quick_c4_32.txt 2>&1
timeout 300 python tools/quick_time.py 640 480 200 256 30 > $O/quick_c4_256.txt 2>&1
tail -3 $O/pytest_gpu.log; cat $O/dd_probe.txt $O/downdate_sweep.txt $O/downdate_sweep_n200.txt; for f in $O/quick_*.txt; do echo $f; tail -2 $f | cut -c1-420; done
