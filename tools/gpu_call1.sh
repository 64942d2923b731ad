#!/bin/bash
# round-2 first GPU call: tests, per-kernel-class times of both operators, ncu --set full captures of the PCG kernels, fp64 FMA rate
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/c1_smi.txt
tools/dfma_bench.bin > gpurun_out/c1_dfma.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c1_pytest.txt 2>&1
echo "pytest rc $?" >> gpurun_out/c1_pytest.txt
for op in assembled matfree; do
RVE_KTIME=1 python tools/profile_solve.py --operator $op --ns 1000 --unique 64 --reps 2 2> gpurun_out/c1_ktime_c2_$op.txt
RVE_KTIME=1 python tools/profile_solve.py --operator $op --workload C4_tangents_10k_reduced_lin --ns 4000 --unique 128 --reps 2 2> gpurun_out/c1_ktime_c4_$op.txt
RVE_KTIME=1 python tools/profile_solve.py --operator $op --workload C3_tangents_10k_reduced_dnn --ns 10000 --unique 128 --reps 2 2> gpurun_out/c1_ktime_c3_$op.txt
done
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:'k_vc_mid|k_direction|k_spmv|k_update' --launch-skip 41 --launch-count 4 -o gpurun_out/c1_ncu_c2 -f \
   python tools/profile_solve.py --operator assembled --ns 250 --unique 32 > gpurun_out/c1_ncu_c2.log 2>&1
$NCU -k regex:'k_vc_mid|k_direction|k_spmv|k_update' --launch-skip 41 --launch-count 4 -o gpurun_out/c1_ncu_c4 -f \
   python tools/profile_solve.py --operator assembled --workload C4_tangents_10k_reduced_lin --ns 2000 --unique 64 > gpurun_out/c1_ncu_c4.log 2>&1
$NCU -k regex:'k_apply' --launch-skip 10 --launch-count 1 -o gpurun_out/c1_ncu_c2mf -f \
   python tools/profile_solve.py --ns 250 --unique 32 > gpurun_out/c1_ncu_c2mf.log 2>&1
$NCU -k regex:'k_apply' --launch-skip 10 --launch-count 1 -o gpurun_out/c1_ncu_c4mf -f \
   python tools/profile_solve.py --workload C4_tangents_10k_reduced_lin --ns 2000 --unique 64 > gpurun_out/c1_ncu_c4mf.log 2>&1
$NCU -k regex:'k_vc_down1|k_vc_up1|k_vc_prolong' --launch-skip 30 --launch-count 3 -o gpurun_out/c1_ncu_c4vc -f \
   python tools/profile_solve.py --workload C4_tangents_10k_reduced_lin --ns 2000 --unique 64 > gpurun_out/c1_ncu_c4vc.log 2>&1
ls -la gpurun_out/c1_* 
