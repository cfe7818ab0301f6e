#!/bin/bash
PFB200_LIB=build_variants/dbg/libpfb200.so timeout 300 python -m pytest tests/test_gpu_configs.py -q -s -k "shard and systematic and 74" 2>&1 | grep -E "abort|passed|failed" | sort | uniq -c | sort -rn | head -30
