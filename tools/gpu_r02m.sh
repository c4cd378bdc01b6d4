#!/bin/bash
O=gpurun_out/r02m
mkdir -p $O
timeout 300 python tools/dd_probe.py 500 72 640 > $O/dd_probe.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lanes or small_update or numeric" > $O/pytest_new.log 2>&1
cat $O/dd_probe.txt; tail -3 $O/pytest_new.log
