#!/bin/bash
# ncu --set full of ONE launch of a kernel + its source page (per-line stall samples); $1 = tag, $2 = kernel regex, $3 = skip, rest = profile_solve args
T=$1; RX=$2; SKIP=$3; shift 3
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"$RX" --launch-skip $SKIP --launch-count 1 -o /tmp/$T -f python tools/profile_solve.py "$@" > gpurun_out/$T.log 2>&1
ncu -i /tmp/$T.ncu-rep --page raw --csv > gpurun_out/${T}_raw.csv 2>/dev/null
ncu -i /tmp/$T.ncu-rep --page source --csv > gpurun_out/${T}_source.csv 2>/dev/null
python tools/ncu_summary.py /tmp/$T.ncu-rep gpurun_out/${T}_summary.csv
ls -la gpurun_out/${T}*
