#!/bin/bash
# round 2, call Q: final-code evidence on one B200: full GPU suite, smoke, bench lines, launch list, full ncu capture of a frame
O=gpurun_out/r02q
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > $O/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > $O/pytest_gpu.log 2>&1; echo "pytest rc $?" >> $O/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1
timeout 600 python bench.py > $O/bench_c3.json 2> $O/bench_c3.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 3 > $O/bench_reference_c3.json 2> $O/bench_reference_c3.err
timeout 400 python bench.py --workload c3full --no-c4-leg > $O/bench_c3full.json 2> $O/bench_c3full.err
timeout 200 python bench.py --workload c2 > $O/bench_c2.json 2> $O/bench_c2.err
timeout 200 python tools/quick_time.py 640 480 500 1 80 > $O/quick_c3.txt 2>&1
timeout 300 python tools/quick_time.py 640 480 200 256 30 > $O/quick_c4_256.txt 2>&1
timeout 120 tools/vendor_bar 3013 640 1 3013 72 1 3013 1000 1 > $O/vendor_bar.txt 2>&1
timeout 200 python tools/downdate_sweep.py 500 72 640 1000 > $O/downdate_sweep_500.txt 2>&1
timeout 600 python tests/report_s3.py $O/s3_config1.json > $O/report_s3.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/launches_bench_c3.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-c4-leg --filter-warm 4 > $O/ncu_bench.log 2>&1
python tools/agg_launches.py $O/launches_bench_c3.csv > $O/launches_bench_c3.txt 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --launch-skip 220 --launch-count 20 -f -o $O/prof_frame_c3 python tools/quick_time.py 640 480 500 1 30 > $O/ncu_frame.log 2>&1
tail -4 $O/pytest_gpu.log; tail -1 $O/smoke.log; cut -c1-1200 $O/bench_c3.json; tail -2 $O/bench_c3.err; cut -c1-600 $O/bench_c3full.json; cut -c1-400 $O/bench_c2.json; cut -c1-500 $O/bench_reference_c3.json; tail -2 $O/quick_c3.txt | cut -c1-500; cat $O/vendor_bar.txt | tail -12; tail -3 $O/report_s3.log | cut -c1-600; head -12 $O/launches_bench_c3.txt; tail -3 $O/ncu_frame.log
