#!/bin/bash
for args in "--N 1000 --opt cluster=8" "--N 1000 --opt cluster=8 --no-smoother" "--N 2000 --opt cluster=16" "--N 4096 --opt cluster=16" "--N 16384 --opt cluster=16" "--N 1000 --opt cluster=8 --model sv" "--N 1000 --opt cluster=8 --dtype float32"; do
  timeout 120 python tools/run_once.py --T 300 $args 2>&1 | tail -1
done
timeout 900 python -m pytest tests/test_gpu_cluster.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -6
exit 0
for args in "--N 1000 --opt cluster=8"; do
  echo "== $args"; timeout 120 python tools/run_once.py --T 300 $args 2>&1 | tail -19 | grep -v "  -  "
done
