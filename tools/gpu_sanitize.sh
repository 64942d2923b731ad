#!/bin/bash
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python tools/profile_solve.py --workload C4_tangents_10k_reduced_lin --ns 4 --unique 4 > gpurun_out/san.txt 2>&1
tail -60 gpurun_out/san.txt
