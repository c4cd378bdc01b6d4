#!/bin/bash
O=gpurun_out/r02ae
mkdir -p $O
timeout 600 python -m pytest tests/test_image_generator.py tests/test_frontend.py tests/test_host_ekf.py -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc $?" >> $O/pytest.log
tail -15 $O/pytest.log
