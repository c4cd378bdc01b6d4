#!/bin/bash
O=gpurun_out/r02l
mkdir -p $O
for cfg in "11=2,12=1" "11=2,2=0" "11=1,2=0" "11=3,12=1"; do
n=$(echo $cfg | tr ',=' '__')
EKFB_OPTS="$cfg" timeout 200 python tools/quick_time.py 640 480 200 32 40 > $O/quick_c4_32_$n.txt 2>&1
done
for cfg in "11=2,12=1" "11=2,2=0"; do
n=$(echo $cfg | tr ',=' '__')
EKFB_OPTS="$cfg" timeout 300 python tools/quick_time.py 640 480 200 256 30 > $O/quick_c4_256_$n.txt 2>&1
done
for f in $O/quick_*.txt; do echo $f; tail -2 $f | head -1 | cut -c1-200; done
