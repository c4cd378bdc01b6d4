#!/bin/bash
# round 2, call B: new bench.py (C4 leg, parity, medians), one-launch chain variant, TMA probe log, ncu of the small kernels
O=gpurun_out/r02b
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q -k "not c3_against and not c4_shape and not s3_sequence_follows" > $O/pytest_gpu.log 2>&1; echo "pytest rc $?" >> $O/pytest_gpu.log
timeout 600 python bench.py > $O/bench_c3.json 2> $O/bench_c3.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 20 > $O/bench_reference_c3.json 2> $O/bench_reference_c3.err
for v in "3=0" "3=3"; do
  EKFB_OPTS=$v timeout 200 python tools/quick_time.py 640 480 500 1 80 > $O/quick_c3_opt${v/=/_}.txt 2>&1
  EKFB_OPTS=$v timeout 200 python tools/quick_time.py 640 480 200 32 30 > $O/quick_c4_32_opt${v/=/_}.txt 2>&1
  EKFB_OPTS=$v timeout 200 python tools/quick_time.py 640 480 200 256 30 > $O/quick_c4_256_opt${v/=/_}.txt 2>&1
  EKFB_OPTS=$v timeout 200 python tools/quick_time.py 320 240 50 1 80 > $O/quick_c2_opt${v/=/_}.txt 2>&1
done
# TMA tensor-map probe: every variant, plain and under compute-sanitizer (the kept log VERDICT r01 asked for)
( for v in 0 1 2 3 8 16 18 4; do echo "== tma_probe $v"; timeout 60 tools/tma_probe $v 2>&1 | grep -v "desc\["; done
  echo "== compute-sanitizer tma_probe 0"; timeout 120 compute-sanitizer tools/tma_probe 0 2>&1 | tail -40
  echo "== compute-sanitizer tma_probe 16"; timeout 120 compute-sanitizer tools/tma_probe 16 2>&1 | tail -40 ) > $O/tma_probe.txt 2>&1
# ncu: one steady-state C3 frame, every kernel, full set (no source import: small report)
timeout 900 ncu --set full --clock-control none --launch-skip 400 --launch-count 40 -f -o $O/prof_frame_c3 python tools/quick_time.py 640 480 500 1 30 > $O/ncu_frame.log 2>&1
timeout 300 ncu --set full --clock-control none -k regex:k_search_ncc --launch-count 2 -f -o $O/prof_ncc python tests/bench_ncc.py 500 > $O/ncu_ncc.log 2>&1
tail -3 $O/pytest_gpu.log; cat $O/bench_c3.json | cut -c1-3000; tail -3 $O/bench_c3.err; cat $O/bench_reference_c3.json | cut -c1-800; for f in $O/quick_*; do echo $f; tail -2 $f | cut -c1-700; done; cat $O/tma_probe.txt | tail -60; ls -la $O
