#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c5_pytest.txt 2>&1
echo "pytest rc $?" >> gpurun_out/c5_pytest.txt
for pl in 1 0; do
RVE_PIPELINE=$pl RVE_KTIME=1 python tools/profile_solve.py --ns 1000 --unique 64 --reps 2 2> gpurun_out/c5_ktime_c2_pl$pl.txt
RVE_PIPELINE=$pl RVE_KTIME=1 python tools/profile_solve.py --workload C3_tangents_10k_reduced_dnn --ns 10000 --unique 128 --reps 2 2> gpurun_out/c5_ktime_c3_pl$pl.txt
done
