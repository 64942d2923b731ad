#!/bin/bash
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:'k_vc_|k_direction|k_apply|k_update|k_fin' --launch-skip 150 --launch-count 15 -o gpurun_out/c6_ncu_c2 -f \
   python tools/profile_solve.py --ns 250 --unique 32 > gpurun_out/c6_ncu_c2.log 2>&1
$NCU -k regex:'k_vc_|k_direction|k_apply|k_update|k_fin' --launch-skip 100 --launch-count 10 -o gpurun_out/c6_ncu_c3 -f \
   python tools/profile_solve.py --workload C3_tangents_10k_reduced_dnn --ns 2500 --unique 64 > gpurun_out/c6_ncu_c3.log 2>&1
