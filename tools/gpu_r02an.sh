#!/bin/bash
O=gpurun_out/r02an
mkdir -p $O
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_ncc.py tests/test_image_generator.py -m gpu -x -q -k "generic_factorisation or dense_slab or single_buffered or tensor_map or image_sequence or phase_by_phase_after" > $O/sanitizer_memcheck.log 2>&1; echo "memcheck rc $?" >> $O/sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "generic_factorisation" > $O/sanitizer_racecheck.log 2>&1; echo "racecheck rc $?" >> $O/sanitizer_racecheck.log
for f in $O/sanitizer_*.log; do echo $f; tail -4 $f; done
