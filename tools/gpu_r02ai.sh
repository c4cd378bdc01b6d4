#!/bin/bash
O=gpurun_out/r02ai
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_map.py tests/test_gpu_edges.py tests/test_golden_frames.py -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc $?" >> $O/pytest.log
timeout 200 python tools/quick_time.py 640 480 500 1 80 > $O/quick_c3.txt 2>&1
timeout 200 python tools/quick_time.py 320 240 50 1 80 > $O/quick_c2.txt 2>&1
tail -3 $O/pytest.log; tail -2 $O/quick_c3.txt | cut -c1-420; tail -1 $O/quick_c2.txt | cut -c1-300
