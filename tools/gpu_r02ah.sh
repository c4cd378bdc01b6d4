#!/bin/bash
O=gpurun_out/r02ah
mkdir -p $O
timeout 300 python tools/quick_time.py 1280 720 1000 1 30 > $O/quick_c5_1000.txt 2>&1
git stash -q 2>/dev/null
tail -2 $O/quick_c5_1000.txt | cut -c1-700
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 700 -c 120 --csv --log-file $O/launches_c5_1000.csv python tools/quick_time.py 1280 720 1000 1 30 > $O/ncu.log 2>&1
python tools/agg_launches.py $O/launches_c5_1000.csv | head -24
