#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_matfree.py -x -q > gpurun_out/c2_pytest.txt 2>&1
echo "pytest rc $?" >> gpurun_out/c2_pytest.txt
RVE_KTIME=1 python tools/profile_solve.py --ns 1000 --unique 64 --reps 2 2> gpurun_out/c2_ktime_c2_matfree.txt
RVE_KTIME=1 python tools/profile_solve.py --workload C4_tangents_10k_reduced_lin --ns 4000 --unique 128 --reps 2 2> gpurun_out/c2_ktime_c4_matfree.txt
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:'k_apply' --launch-skip 10 --launch-count 1 -o gpurun_out/c2_ncu_c2mf -f \
   python tools/profile_solve.py --ns 250 --unique 32 > gpurun_out/c2_ncu_c2mf.log 2>&1
