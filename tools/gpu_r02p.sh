#!/bin/bash
O=gpurun_out/r02p
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py tests/test_gpu_map.py tests/test_parity_bench_sizes.py tests/test_host_ekf.py -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc $?" >> $O/pytest_gpu.log
timeout 200 python tools/quick_time.py 640 480 500 1 80 > $O/quick_c3.txt 2>&1
timeout 200 python tools/quick_time.py 320 240 50 1 80 > $O/quick_c2.txt 2>&1
timeout 200 python tools/quick_time.py 640 480 200 1 60 > $O/quick_n200.txt 2>&1
timeout 200 python tools/quick_time.py 640 480 200 32 40 > $O/quick_c4_32.txt 2>&1
timeout 300 python tools/quick_time.py 640 480 200 256 30 > $O/quick_c4_256.txt 2>&1
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "phase_by_phase_c2 or small_update_kernel or lanes_give or whole_step_sequence" > $O/sanitizer_memcheck.log 2>&1; echo "memcheck rc $?" >> $O/sanitizer_memcheck.log
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "phase_by_phase_c2 or small_update_kernel" > $O/sanitizer_racecheck.log 2>&1; echo "racecheck rc $?" >> $O/sanitizer_racecheck.log
tail -3 $O/pytest_gpu.log; for f in $O/quick_*.txt; do echo $f; tail -2 $f | cut -c1-420; done; tail -5 $O/sanitizer_memcheck.log; tail -5 $O/sanitizer_racecheck.log
