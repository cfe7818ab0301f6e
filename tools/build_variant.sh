#!/bin/bash
# Builds a tuning / profiling variant of libpfb200.so into build_variants/NAME (git-ignored; selected at
# run time with PFB200_LIB=build_variants/NAME/libpfb200.so).  Usage:
#   tools/build_variant.sh NAME [lgss|all] [extra nvcc flags, e.g. -DPFB_PROFILE -DPFB_CHASE=21]
set -e
NAME=$1; shift
WHAT=${1:-lgss}; shift || true
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=$ROOT/build_variants/$NAME
mkdir -p $OUT
if [ "$WHAT" = lgss ]; then MODELS="0"; else MODELS="0 1 2 3"; fi
make -s -j8 -C $ROOT/qnmh_sysid2018_b200/csrc OUT=$OUT MODELS="$MODELS" EXTRA="$*" 2>&1 | grep -E "error|spill|registers" || true
ls -la $OUT/libpfb200.so
