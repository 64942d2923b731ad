#!/bin/bash
# bench lines of the three main workloads (tag = $1)
T=${1:-x}
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/${T}_bench_C2.json 2> gpurun_out/${T}_bench_C2.err
python bench.py --steps 5 --warmup 3 --workload C3_tangents_10k_reduced_dnn --no-cpu-baseline > gpurun_out/${T}_bench_C3.json 2> gpurun_out/${T}_bench_C3.err
python bench.py --steps 5 --warmup 3 --workload C4_tangents_10k_reduced_lin --ns 4000 --no-cpu-baseline > gpurun_out/${T}_bench_C4lin.json 2> gpurun_out/${T}_bench_C4lin.err
tail -c 600 gpurun_out/${T}_bench_C2.err
