#!/bin/bash
# ncu --set full captures of ONE PCG iteration (C2 K=2 at 250 full RVEs, C3 K=3 at 2500 reduced RVEs) + whole-run launch lists; tag = $1
T=${1:-r02j}
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
cap() {   # name, kernel regex, skip, count, profile_solve args...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  $NCU -k regex:"$rx" --launch-skip $skip --launch-count $cnt -o /tmp/${name} -f python tools/profile_solve.py "$@" > gpurun_out/${name}.log 2>&1
  ncu -i /tmp/${name}.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
  python tools/ncu_summary.py /tmp/${name}.ncu-rep gpurun_out/${name}_summary.csv
}
cap ${T}_ncu_c2 'k_vc_|k_direction|k_apply|k_update|k_fin' 168 14 --ns 250 --unique 32
cap ${T}_ncu_c3 'k_vc_|k_direction|k_apply|k_update|k_fin' 110 10 --workload C3_tangents_10k_reduced_dnn --ns 2500 --unique 64
for w in c2 c3; do
  args="--ns 250 --unique 32"; [ $w = c3 ] && args="--workload C3_tangents_10k_reduced_dnn --ns 2500 --unique 64"
  ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${T}_launches_${w}_raw.csv python tools/profile_solve.py $args > gpurun_out/${T}_ll_${w}.log 2>&1
  python tools/launch_summary.py gpurun_out/${T}_launches_${w}_raw.csv gpurun_out/${T}_launches_${w}_summary.csv \
    "# ${T}: ncu --metrics gpu__time_duration.sum --clock-control none: python tools/profile_solve.py $args (whole run of one chunk: set_batch, assemble, solve, epilogue)" \
    "# cold-cache serialised per-launch times: compare SHARES"
done
ls -la gpurun_out/${T}_*summary.csv
