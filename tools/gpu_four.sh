#!/bin/bash
NG=$(nvidia-smi -L | wc -l)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus $NG --steps 3 --warmup 3 > gpurun_out/r2_x${NG}_bench.json 2> gpurun_out/r2_x${NG}_bench.err
tail -c 200 gpurun_out/r2_x${NG}_bench.json; tail -2 gpurun_out/r2_x${NG}_bench.err | cut -c1-200
