#!/bin/bash
echo "== default"; python tools/run_once.py --T 100 2>&1 | grep "kernel ms"; python tools/run_once.py --T 100 --N 4096 --B 1024 2>&1 | grep "kernel ms"
for so in build_variants/*.so; do
  echo "== $so"
  PFB200_LIB=$PWD/$so python tools/run_once.py --T 100 2>&1 | grep -E "kernel ms"
  PFB200_LIB=$PWD/$so python tools/run_once.py --T 100 --N 4096 --B 1024 2>&1 | grep -E "kernel ms"
done
