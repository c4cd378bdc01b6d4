#!/bin/bash
O=gpurun_out/r02al
mkdir -p $O
timeout 300 python tools/quick_time.py 1280 720 1000 1 30 > $O/quick_c5_1000_global.txt 2>&1
EKFB_OPTS="17=1152" timeout 300 python tools/quick_time.py 1280 720 1000 1 30 > $O/quick_c5_1000_slab.txt 2>&1
for f in $O/quick_*.txt; do echo $f; tail -2 $f | cut -c1-330; done
