#!/bin/bash
mkdir -p gpurun_out
RVE_DEBUG_SYNC=1 timeout 1200 python -m pytest tests -m gpu -x -q -k "$1" > gpurun_out/gt_pytest.txt 2>&1
echo "pytest rc $?" >> gpurun_out/gt_pytest.txt
tail -n 25 gpurun_out/gt_pytest.txt
