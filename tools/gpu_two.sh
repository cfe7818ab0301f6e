#!/bin/bash
# 2 GPUs: the distributed-filter parity test, then the driver's multi-GPU bench command
nvidia-smi --query-gpu=name --format=csv,noheader | head -3
timeout 600 python -m pytest tests/test_gpu_distributed.py -m gpu -q 2>&1 | tail -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_x2_bench.json 2> gpurun_out/r2_x2_bench.err
tail -c 3000 gpurun_out/r2_x2_bench.json; tail -3 gpurun_out/r2_x2_bench.err
