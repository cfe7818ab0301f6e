#!/bin/bash
# round-2 third session: A/B of instruction-diet variants (C2 fp64 / fp32, C3) -- kernel time per variant
mkdir -p gpurun_out
for v in "$@"; do
  echo "== variant $v"
  export PFB200_LIB=build_variants/$v/libpfb200.so
  timeout 200 python tools/run_once.py --T 300 2>&1 | tail -1
  timeout 200 python tools/run_once.py --T 300 --dtype float32 2>&1 | tail -1
  timeout 200 python tools/run_once.py --T 100 --N 4096 --B 1024 2>&1 | tail -1
done
