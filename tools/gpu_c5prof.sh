#!/bin/bash
NG=$(nvidia-smi -L | wc -l)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1"
export PFB200_LIB=build_variants/h_prof/libpfb200.so
timeout 300 $TR --master-port 29551 tools/run_dist.py --log2 24 --T 50 2>&1 | grep -E "C5|us  max|total|rror" | tail -20
unset PFB200_LIB
timeout 300 $TR --master-port 29552 tools/run_dist.py --log2 24 --T 50 --no-smoother 2>&1 | grep -E "C5|rror" | tail -3
timeout 300 $TR --master-port 29553 tools/run_dist.py --log2 24 --T 50 --L 1 2>&1 | grep -E "C5|rror" | tail -3
timeout 100 python tools/run_once.py --N 2097152 --T 50 --G 128 2>&1 | tail -1
timeout 100 python tools/run_once.py --N 2097152 --T 50 --G 128 --inject 2>&1 | tail -1
