#!/bin/bash
# per-phase cycle counters (-DPFB_PROFILE build) of the latency-bound configurations
export PFB200_LIB=$PWD/build_variants/prof/libpfb200.so
echo "== C1 N=1000";       timeout 200 python tools/run_once.py --T 300 --N 1000 2>&1 | tail -19
echo "== C1 N=1000 no smoother"; timeout 200 python tools/run_once.py --T 300 --N 1000 --no-smoother 2>&1 | tail -19
echo "== C4 sv N=65536";   timeout 200 python tools/run_once.py --T 300 --N 65536 --model sv 2>&1 | tail -19
echo "== C4 sv N=65536 no smoother";   timeout 200 python tools/run_once.py --T 300 --N 65536 --model sv --no-smoother 2>&1 | tail -19
echo "== lgss N=65536 G=32";   timeout 200 python tools/run_once.py --T 300 --N 65536 --G 32 2>&1 | tail -19
echo "== lgss N=65536 G=16";   timeout 200 python tools/run_once.py --T 300 --N 65536 --G 16 2>&1 | tail -19
echo "== C2"; timeout 200 python tools/run_once.py --T 300 2>&1 | tail -19
