#!/bin/bash
# 8 GPUs: the driver's multi-GPU bench command (independent C2 replicates + the c3_strong / c5 records)
nvidia-smi --query-gpu=name --format=csv,noheader | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2_x8_bench.json 2> gpurun_out/r2_x8_bench.err
tail -c 2500 gpurun_out/r2_x8_bench.json; tail -2 gpurun_out/r2_x8_bench.err
