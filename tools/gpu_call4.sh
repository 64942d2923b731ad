#!/bin/bash
mkdir -p gpurun_out
RVE_KTIME=1 RVE_KTIME_SERIES=1 python tools/profile_solve.py --ns 1000 --unique 256 --reps 2 2> gpurun_out/c4_series_c2.txt
RVE_KTIME=1 RVE_KTIME_SERIES=1 python tools/profile_solve.py --workload C3_tangents_10k_reduced_dnn --ns 10000 --unique 256 --reps 2 2> gpurun_out/c4_series_c3.txt
