#!/bin/bash
# 8 GPUs of one box: BASELINE config 5 as a real campaign (files in, files out) and the default C2 bench line
PER=${PER:-4000}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/g8_smi.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/run_c5_campaign.py --per-gpu $PER > gpurun_out/g8_c5_campaign.log 2>&1
echo "campaign rc $?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/g8_bench_C2_8gpu.json 2> gpurun_out/g8_bench_C2_8gpu.err
echo "bench rc $?"
tail -n 3 gpurun_out/g8_c5_campaign.log | cut -c1-900
