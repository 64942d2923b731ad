#!/bin/bash
# whole-run launch lists (set_batch -> assemble -> solve -> epilogue) of one C2 chunk (250 full RVEs) and one C3 chunk (2500 reduced)
T=${1:-r02f}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${T}_launches_c2_raw.csv \
    python tools/profile_solve.py --ns 250 --unique 32 > gpurun_out/${T}_ll_c2.log 2>&1
python tools/launch_summary.py gpurun_out/${T}_launches_c2_raw.csv gpurun_out/${T}_launches_c2_summary.csv \
  "# ${T}: ncu --metrics gpu__time_duration.sum --clock-control none: python tools/profile_solve.py --ns 250 --unique 32 (C2 chunk: 250 full RVEs, per, K=2; whole run: set_batch, assemble, solve, snapshot epilogue)" \
  "# cold-cache serialised per-launch times: compare SHARES"
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${T}_launches_c3_raw.csv \
    python tools/profile_solve.py --workload C3_tangents_10k_reduced_dnn --ns 2500 --unique 64 > gpurun_out/${T}_ll_c3.log 2>&1
python tools/launch_summary.py gpurun_out/${T}_launches_c3_raw.csv gpurun_out/${T}_launches_c3_summary.csv \
  "# ${T}: same for python tools/profile_solve.py --workload C3_tangents_10k_reduced_dnn --ns 2500 --unique 64 (C3 chunk: 2500 reduced RVEs, dnn, K=3)" \
  "# cold-cache serialised per-launch times: compare SHARES"
rm -f gpurun_out/${T}_launches_c?_raw.csv.tmp
head -40 gpurun_out/${T}_launches_c2_summary.csv
