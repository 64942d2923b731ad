#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/bis3.txt
for d in 9 10 11 12; do
echo "== dbg $d" >> gpurun_out/bis3.txt
RVE_SYM_DBG=$d RVE_DEBUG_SYNC=1 timeout 300 python tools/profile_solve.py --operator assembled --workload C4_tangents_10k_reduced_lin --ns 4 --unique 4 2>&1 | tail -n 1 >> gpurun_out/bis3.txt
done
cat gpurun_out/bis3.txt
