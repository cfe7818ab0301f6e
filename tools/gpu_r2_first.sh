#!/bin/bash
# round 2, first GPU call: the whole GPU test tier (old + new configurations), then per-phase cycle
# counters of the profiling build for C2 / C4 / C1 / C3
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader
timeout 1500 python -m pytest tests -m gpu -q -x -s 2>&1 | grep -v "^$" | tail -40
export PFB200_LIB=$PWD/build_variants/prof/libpfb200.so
echo "== profile C2 f64"; timeout 200 python tools/run_once.py --T 300 2>&1 | tail -19
echo "== profile C2 f32"; timeout 200 python tools/run_once.py --T 300 --dtype float32 2>&1 | tail -19
echo "== profile C4 sv";  timeout 200 python tools/run_once.py --T 300 --N 65536 --model sv 2>&1 | tail -19
echo "== profile C1";     timeout 200 python tools/run_once.py --T 300 --N 1000 2>&1 | tail -19
echo "== profile C3";     timeout 200 python tools/run_once.py --T 100 --N 4096 --B 1024 2>&1 | tail -19
