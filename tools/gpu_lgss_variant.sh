#!/bin/bash
# LGSS-only tuning build: parity tests that only need the LGSS kernels, then timings
V=${1:-base}
export PFB200_LIB=$PWD/build_variants/$V/libpfb200.so
timeout 900 python -m pytest tests -m gpu -q -x -k "lgss or geometries or stratified or multinomial_resampling or multinomial_philox or fixed_initial or batched or ring_mode or philox_path or device_normals or against_exact_kalman or philox_loglik or c2_geometry or c3_inst or fp32 or sharded or integration or dropin_filter" 2>&1 | tail -5
timeout 120 python tools/run_once.py --T 300 2>&1 | tail -1
timeout 120 python tools/run_once.py --T 300 --dtype float32 2>&1 | tail -1
timeout 120 python tools/run_once.py --T 100 --N 4096 --B 4096 2>&1 | tail -1
timeout 120 python tools/run_once.py --T 300 --N 1000 2>&1 | tail -1
timeout 120 python tools/run_once.py --T 300 --N 65536 2>&1 | tail -1
timeout 120 python tools/run_once.py --T 50 --N 16777216 2>&1 | tail -1
