#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "matfree or assembly or parity" > gpurun_out/gt_pytest.txt 2>&1
echo "pytest rc $?" >> gpurun_out/gt_pytest.txt
python bench.py --steps 3 --warmup 3 --workload C3_tangents_10k_reduced_dnn --operator assembled --no-cpu-baseline --no-extra > gpurun_out/r02e_bench_C3_assembled.json 2> gpurun_out/r02e_bench_C3_assembled.err
python bench.py --steps 3 --warmup 3 --workload C4_tangents_10k_reduced_lin --ns 4000 --operator assembled --no-cpu-baseline --no-extra > gpurun_out/r02e_bench_C4lin_assembled.json 2> gpurun_out/r02e_bench_C4lin_assembled.err
python bench.py --steps 3 --warmup 3 --operator assembled --no-cpu-baseline --no-extra > gpurun_out/r02e_bench_C2_assembled.json 2> gpurun_out/r02e_bench_C2_assembled.err
ncu --set full --clock-control none --import-source on -k regex:'k_spmv' --launch-skip 10 --launch-count 1 -o /tmp/sp3 -f python tools/profile_solve.py --operator assembled --workload C3_tangents_10k_reduced_dnn --ns 2500 --unique 64 > gpurun_out/r02e_ncu_spmv_c3.log 2>&1
ncu -i /tmp/sp3.ncu-rep --page raw --csv > gpurun_out/r02e_ncu_spmv_c3_raw.csv 2>/dev/null
python tools/ncu_summary.py /tmp/sp3.ncu-rep gpurun_out/r02e_ncu_spmv_c3_summary.csv
tail -n 3 gpurun_out/gt_pytest.txt
