#!/bin/bash
NG=$(nvidia-smi -L | wc -l)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
timeout 300 python -m pytest tests/test_gpu_distributed.py -q 2>&1 | tail -1
timeout 300 python -m pytest tests/test_gpu_configs.py -q -k "shard or dist" 2>&1 | tail -1
timeout 300 $TR --master-port 29551 tools/run_dist.py --log2 24 26 --T 50 2>&1 | grep -E "C5|rror" | tail -4
if [ $NG -ge 4 ]; then timeout 300 $TR --master-port 29552 tools/run_dist.py --log2 24 26 --T 50 --ctas 148 2>&1 | grep -E "C5|rror" | tail -4; fi
