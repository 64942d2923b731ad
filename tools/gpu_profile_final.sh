#!/bin/bash
# evidence for profiles/: launch list of a short bench run + ncu --set full captures of one PCG iteration (C2 K=2, C3 K=3);
# the .ncu-rep files are reduced to CSV on the box (gpurun returns at most 64 MiB)
T=${1:-r02}
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/${T}_plain_bench.json 2> gpurun_out/${T}_plain_bench.err
echo "plain rc $?"
ncu --metrics gpu__time_duration.sum --clock-control none -s 2600 -c 700 --csv --log-file gpurun_out/${T}_launches_raw.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra > gpurun_out/${T}_ncu_list.log 2>&1
NCU="ncu --set full --clock-control none --import-source on"
cap() {   # name, kernel regex, skip, count, profile_solve args...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  $NCU -k regex:"$rx" --launch-skip $skip --launch-count $cnt -o /tmp/${name} -f python tools/profile_solve.py "$@" > gpurun_out/${name}.log 2>&1
  ncu -i /tmp/${name}.ncu-rep --page raw --csv > gpurun_out/${name}_raw.csv 2>/dev/null
  python tools/ncu_summary.py /tmp/${name}.ncu-rep gpurun_out/${name}_summary.csv
}
cap ${T}_ncu_c2 'k_vc_|k_direction|k_apply|k_update|k_fin' 160 16 --ns 250 --unique 32
cap ${T}_ncu_c3 'k_vc_|k_direction|k_apply|k_update|k_fin' 110 11 --workload C3_tangents_10k_reduced_dnn --ns 2500 --unique 64
cap ${T}_ncu_spmv_c2 'k_spmv' 10 1 --operator assembled --ns 250 --unique 32
cap ${T}_ncu_spmv_c3 'k_spmv' 10 1 --operator assembled --workload C3_tangents_10k_reduced_dnn --ns 2500 --unique 64
cp /tmp/${T}_ncu_spmv_c2.ncu-rep gpurun_out/ 2>/dev/null
du -sh gpurun_out; ls -la gpurun_out/${T}_*
