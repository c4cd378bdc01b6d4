#!/bin/bash
# round 2, call A: the new parity tests at the benchmarked sizes, the vendor bar, chain probes
O=gpurun_out/r02a
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > $O/smi.txt 2>&1
nproc > $O/nproc.txt
timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 -s > $O/pytest_gpu.log 2>&1; echo "pytest rc $?" >> $O/pytest_gpu.log
timeout 120 tools/vendor_bar > $O/vendor_bar.txt 2>&1
timeout 200 python tools/downdate_sweep.py 500 72 640 1000 > $O/downdate_sweep_500.txt 2>&1
timeout 200 python tools/dbg_cycles.py > $O/dbg_cycles.txt 2>&1
timeout 300 python tools/quick_time.py 640 480 500 1 80 > $O/quick_c3.txt 2>&1
tail -5 $O/pytest_gpu.log; cat $O/vendor_bar.txt; cat $O/downdate_sweep_500.txt; cat $O/dbg_cycles.txt; tail -2 $O/quick_c3.txt
