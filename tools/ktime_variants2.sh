#!/bin/bash
# RVE_KTIME per-kernel-class times of C2 (1000 full RVEs) and C3 (10k reduced) for the main library and every variant
mkdir -p gpurun_out
for lib in deepbnd_b200/lib/librve_b200.so deepbnd_b200/lib/variants/*.so; do
  [ -f "$lib" ] || continue
  n=$(basename $lib .so)
  DEEPBND_RVE_LIB=$PWD/$lib RVE_DEBUG_TIMING=1 RVE_KTIME=1 python tools/profile_solve.py --ns 1000 --unique 64 --reps 2 2> gpurun_out/kv_${n}_c2.txt
  [ -n "$SKIP_C3" ] || DEEPBND_RVE_LIB=$PWD/$lib RVE_DEBUG_TIMING=1 RVE_KTIME=1 python tools/profile_solve.py --workload C3_tangents_10k_reduced_dnn --ns 10000 --unique 128 --reps 2 2> gpurun_out/kv_${n}_c3.txt
  for c in c2 c3; do
    [ -f gpurun_out/kv_${n}_${c}.txt ] && echo "$n $c: $(grep 'per iteration' gpurun_out/kv_${n}_${c}.txt | tail -1 | sed 's/.*per iteration (us)//') | $(tail -1 gpurun_out/kv_${n}_${c}.txt | grep -o "iters mean [0-9.]*") $(tail -1 gpurun_out/kv_${n}_${c}.txt | grep -o "'pcg': [0-9.]*")"
  done
done
