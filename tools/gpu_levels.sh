#!/bin/bash
# iteration count and PCG time against the depth of the coarse hierarchy
for L in 8 5 4; do
  echo "C2 RVE_MAXLEVELS=$L: $(RVE_MAXLEVELS=$L RVE_DEBUG_TIMING=1 RVE_KTIME=1 python tools/profile_solve.py --ns 1000 --unique 64 --reps 2 2>&1 | grep -E 'per iteration|iters mean' | tail -2 | sed 's/.*per iteration (us)//; s/set_batch_numbering.*precond_setup/precond_setup/' | tr '\n' ' ')"
done
for L in 8 3 2; do
  echo "C3 RVE_MAXLEVELS=$L: $(RVE_MAXLEVELS=$L RVE_DEBUG_TIMING=1 RVE_KTIME=1 python tools/profile_solve.py --workload C3_tangents_10k_reduced_dnn --ns 10000 --unique 128 --reps 2 2>&1 | grep -E 'per iteration|iters mean' | tail -2 | sed 's/.*per iteration (us)//; s/set_batch_numbering.*precond_setup/precond_setup/' | tr '\n' ' ')"
done
