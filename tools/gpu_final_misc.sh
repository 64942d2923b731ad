#!/bin/bash
# r02h: C2 with morphed meshes (shared topology), whole-run launch list of one C2 chunk
T=${1:-r02h}
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra --mesher morph > gpurun_out/${T}_bench_C2_morph_shared.json 2> gpurun_out/${T}_bench_C2_morph_shared.err
echo "morph rc $?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${T}_launches_c2_raw.csv \
    python tools/profile_solve.py --ns 250 --unique 32 > gpurun_out/${T}_ll_c2.log 2>&1
python tools/launch_summary.py gpurun_out/${T}_launches_c2_raw.csv gpurun_out/${T}_launches_c2_summary.csv \
  "# ${T}: ncu --metrics gpu__time_duration.sum --clock-control none: python tools/profile_solve.py --ns 250 --unique 32 (C2 chunk: 250 full RVEs, per, K=2; whole run: set_batch, assemble, solve, snapshot epilogue)" \
  "# cold-cache serialised per-launch times: compare SHARES"
head -45 gpurun_out/${T}_launches_c2_summary.csv
