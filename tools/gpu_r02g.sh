#!/bin/bash
O=gpurun_out/r02g
mkdir -p $O
timeout 60 tools/dmma_probe2 > $O/dmma_probe2.txt 2>&1
timeout 200 python tools/dbg_chain.py > $O/dbg_chain.txt 2>&1
timeout 200 python tools/quick_time.py 640 480 500 1 80 > $O/quick_c3.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc $?" >> $O/pytest_gpu.log
cat $O/dmma_probe2.txt $O/dbg_chain.txt; tail -2 $O/quick_c3.txt | cut -c1-800; tail -3 $O/pytest_gpu.log
