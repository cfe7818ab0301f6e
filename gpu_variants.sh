#!/bin/bash
for so in build_variants/*.so; do
  echo "== $so"
  PFB200_LIB=$PWD/$so python tools/run_once.py --T 200 2>&1 | tail -1
  PFB200_LIB=$PWD/$so python tools/run_once.py --T 200 --N 65536 2>&1 | tail -1
  PFB200_LIB=$PWD/$so python tools/run_once.py --T 200 --N 4096 --B 1024 2>&1 | tail -1
done
echo "== default build"; python -m pytest tests -m gpu -q 2>&1 | tail -3
