#!/bin/bash
# 8 GPUs, end of round 2: the sharded-filter parity test, the C5 sweep, then the driver's multi-GPU bench command
NG=$(nvidia-smi -L | wc -l)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
timeout 300 python -m pytest tests/test_gpu_distributed.py -q 2>&1 | tail -1
timeout 300 $TR --master-port 29551 tools/run_dist.py --log2 24 26 --T 50 2>&1 | grep -E "C5|rror" | tail -4
timeout 900 $TR --master-port 29542 bench.py --gpus $NG --steps 5 --warmup 3 > gpurun_out/r2_x${NG}_bench.json 2> gpurun_out/r2_x${NG}_bench.err
tail -c 300 gpurun_out/r2_x${NG}_bench.json; tail -2 gpurun_out/r2_x${NG}_bench.err
