#!/bin/bash
# default bench line (C2 + extra_workloads) and the C5-resolution workload; tag = $1
T=${1:-x}
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/${T}_bench_default.json 2> gpurun_out/${T}_bench_default.err
echo "default rc $?"
python bench.py --steps 3 --warmup 3 --workload C5_snapshots_refined_per --no-cpu-baseline --unique 64 > gpurun_out/${T}_bench_C5.json 2> gpurun_out/${T}_bench_C5.err
echo "C5 rc $?"
tail -c 400 gpurun_out/${T}_bench_default.err gpurun_out/${T}_bench_C5.err
