#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/gt_pytest.txt 2>&1
echo "pytest rc $?" >> gpurun_out/gt_pytest.txt
tail -n 25 gpurun_out/gt_pytest.txt
