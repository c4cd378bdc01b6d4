#!/bin/bash
# round 2, call AB: final code -- full GPU suite, smoke, bench lines
O=gpurun_out/r02ab
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q --durations=5 > $O/pytest_gpu.log 2>&1; echo "pytest rc $?" >> $O/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1
timeout 600 python bench.py > $O/bench_c3.json 2> $O/bench_c3.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 3 > $O/bench_reference_c3.json 2> $O/bench_reference_c3.err
timeout 400 python bench.py --workload c3full --no-c4-leg > $O/bench_c3full.json 2> $O/bench_c3full.err
timeout 200 python bench.py --workload c2 > $O/bench_c2.json 2> $O/bench_c2.err
timeout 200 python tools/quick_time.py 640 480 500 1 80 > $O/quick_c3.txt 2>&1
timeout 200 python tools/quick_time.py 320 240 50 1 80 > $O/quick_c2.txt 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/launches_bench_c3.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-c4-leg --filter-warm 4 > $O/ncu_bench.log 2>&1
python tools/agg_launches.py $O/launches_bench_c3.csv > $O/launches_bench_c3.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --launch-skip 220 --launch-count 20 -f -o $O/prof_frame_c3 python tools/quick_time.py 640 480 500 1 30 > $O/ncu_frame.log 2>&1
tail -3 $O/pytest_gpu.log; tail -1 $O/smoke.log; cut -c1-250 $O/bench_c3.json; tail -2 $O/bench_c3.err; tail -2 $O/quick_c3.txt | cut -c1-400; tail -2 $O/quick_c2.txt | cut -c1-300
