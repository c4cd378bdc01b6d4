#!/bin/bash
# round 2, call C: chain + TRSM in one launch (variant 4), prefetching critical CTA, downdate changes, second TMA probe
O=gpurun_out/r02c
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc $?" >> $O/pytest_gpu.log
EKFB_OPTS="3=4" timeout 900 python -m pytest tests/test_parity_bench_sizes.py tests/test_gpu_parity.py tests/test_golden_frames.py -m gpu -x -q -s -k "c3_against or full_size or whole_step or golden or batched" > $O/pytest_gpu_variant4.log 2>&1; echo "pytest rc $?" >> $O/pytest_gpu_variant4.log
for v in "3=0" "3=3" "3=4" "3=4,2=2" "3=0,2=2"; do
  EKFB_OPTS=$v timeout 200 python tools/quick_time.py 640 480 500 1 80 > $O/quick_c3_opt${v//[=,]/_}.txt 2>&1
  EKFB_OPTS=$v timeout 200 python tools/quick_time.py 320 240 50 1 80 > $O/quick_c2_opt${v//[=,]/_}.txt 2>&1
  EKFB_OPTS=$v timeout 200 python tools/quick_time.py 640 480 200 1 60 > $O/quick_n200_opt${v//[=,]/_}.txt 2>&1
done
timeout 200 python tools/downdate_sweep.py 500 72 640 1000 > $O/downdate_sweep_500.txt 2>&1
timeout 200 python tools/downdate_sweep.py 200 64 290 > $O/downdate_sweep_200.txt 2>&1
( for m in 0 1 2; do echo "== tma_probe2 mode $m"; timeout 60 tools/tma_probe2 $m 2>&1; done
  for b in tma_probe tma_probe_w32 tma_probe_w64; do echo "== $b 0 (round-1 probe, UINT8 box 48 / 32 / 64 bytes wide x 36 rows)"; timeout 60 tools/$b 0 2>&1 | grep -v "desc\["; done
  echo "== compute-sanitizer tma_probe2 0"; timeout 120 compute-sanitizer tools/tma_probe2 0 2>&1 | tail -30 ) > $O/tma_probe2.txt 2>&1
tail -4 $O/pytest_gpu.log; tail -6 $O/pytest_gpu_variant4.log; for f in $O/quick_*; do echo $f; tail -2 $f | cut -c1-700; done; cat $O/downdate_sweep_*.txt; cat $O/tma_probe2.txt
