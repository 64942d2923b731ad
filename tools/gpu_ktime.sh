#!/bin/bash
# per-kernel-class times of C2 (1000 full RVEs) and C3 (10k reduced RVEs); tag = $1, extra env passed through
T=${1:-kt}
mkdir -p gpurun_out
RVE_DEBUG_TIMING=1 RVE_KTIME=1 python tools/profile_solve.py --ns 1000 --unique 64 --reps 2 2> gpurun_out/${T}_c2.txt
RVE_DEBUG_TIMING=1 RVE_KTIME=1 python tools/profile_solve.py --workload C3_tangents_10k_reduced_dnn --ns 10000 --unique 128 --reps 2 2> gpurun_out/${T}_c3.txt
grep "per iteration" gpurun_out/${T}_c2.txt gpurun_out/${T}_c3.txt | tail -n 4
