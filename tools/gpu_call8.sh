#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c8_pytest.txt 2>&1
echo "pytest rc $?" >> gpurun_out/c8_pytest.txt
bash tools/gpu_ktime.sh kt8
