#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -k "shared_topology or morphed" > gpurun_out/gt_pytest.txt 2>&1
echo "pytest rc $?" >> gpurun_out/gt_pytest.txt
bash tools/gpu_e2e.sh r02c
