#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/g8_smi.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/run_c5_campaign.py --per-gpu 2000 > gpurun_out/g8_c5_campaign.log 2>&1
echo "campaign rc $?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/g8_bench_C2_8gpu.json 2> gpurun_out/g8_bench_C2_8gpu.err
echo "bench rc $?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 3 --warmup 3 --workload C5_snapshots_refined_per --unique 64 > gpurun_out/g8_bench_C5_8gpu.json 2> gpurun_out/g8_bench_C5_8gpu.err
echo "bench C5 rc $?"
tail -n 3 gpurun_out/g8_c5_campaign.log | cut -c1-600
