#!/bin/bash
# round 2, call D: new defaults (fused chain + TRSM for a single filter, TMA-fed downdate for all), full GPU suite, bench lines
O=gpurun_out/r02d
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 -s > $O/pytest_gpu.log 2>&1; echo "pytest rc $?" >> $O/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1
timeout 600 python bench.py > $O/bench_c3.json 2> $O/bench_c3.err
timeout 400 python bench.py --workload c3full --no-c4-leg > $O/bench_c3full.json 2> $O/bench_c3full.err
timeout 200 python bench.py --workload c2 --no-cpu-baseline > $O/bench_c2.json 2> $O/bench_c2.err
for v in "2=2" "2=0"; do
  EKFB_OPTS=$v timeout 200 python tools/quick_time.py 640 480 200 32 30 > $O/quick_c4_32_opt${v//[=,]/_}.txt 2>&1
  EKFB_OPTS=$v timeout 200 python tools/quick_time.py 640 480 200 256 30 > $O/quick_c4_256_opt${v//[=,]/_}.txt 2>&1
done
timeout 200 python tools/quick_time.py 640 480 500 1 80 > $O/quick_c3.txt 2>&1
timeout 120 tools/vendor_bar 3013 640 1 3013 72 1 3013 1000 1 > $O/vendor_bar.txt 2>&1
timeout 200 python tools/downdate_sweep.py 500 72 640 1000 > $O/downdate_sweep_500.txt 2>&1
timeout 900 ncu --set full --clock-control none --launch-skip 300 --launch-count 30 -f -o $O/prof_frame_c3 python tools/quick_time.py 640 480 500 1 30 > $O/ncu_frame.log 2>&1
tail -4 $O/pytest_gpu.log; tail -1 $O/smoke.log; cut -c1-1500 $O/bench_c3.json; tail -2 $O/bench_c3.err; cut -c1-900 $O/bench_c3full.json; tail -2 $O/bench_c3full.err; cut -c1-600 $O/bench_c2.json; for f in $O/quick_*; do echo $f; tail -2 $f | cut -c1-600; done; cat $O/vendor_bar.txt $O/downdate_sweep_500.txt
