#!/bin/bash
# e2e of C2 with the three mesh supplies (delaunay / morph per-RVE symbolic / morph shared topology) and C5
mkdir -p gpurun_out
T=${1:-x}
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra --mesher morph --no-shared-topology > gpurun_out/${T}_C2_morph_perrve.json 2> gpurun_out/${T}_C2_morph_perrve.err
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra --mesher morph > gpurun_out/${T}_C2_morph_shared.json 2> gpurun_out/${T}_C2_morph_shared.err
python bench.py --steps 3 --warmup 3 --workload C5_snapshots_refined_per --no-cpu-baseline --unique 64 > gpurun_out/${T}_C5_shared.json 2> gpurun_out/${T}_C5_shared.err
tail -c 300 gpurun_out/${T}_*.err
