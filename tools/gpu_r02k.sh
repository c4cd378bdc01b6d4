#!/bin/bash
O=gpurun_out/r02k
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_parity_bench_sizes.py tests/test_gpu_map.py -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc $?" >> $O/pytest_gpu.log
for L in 1 2 3 4; do
EKFB_OPTS="11=$L" timeout 200 python tools/quick_time.py 640 480 200 32 40 > $O/quick_c4_32_lanes$L.txt 2>&1
done
for L in 1 2 4; do
EKFB_OPTS="11=$L" timeout 300 python tools/quick_time.py 640 480 200 256 30 > $O/quick_c4_256_lanes$L.txt 2>&1
done
tail -3 $O/pytest_gpu.log; for f in $O/quick_*.txt; do echo $f; tail -2 $f | head -1 | cut -c1-200; done
