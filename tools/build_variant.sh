#!/bin/bash
# Builds a tuning / profiling variant of libpfb200.so into build_variants/NAME (git-ignored; selected at
# run time with PFB200_LIB=build_variants/NAME/libpfb200.so).  Usage:
#   tools/build_variant.sh NAME [lgss|all] [extra nvcc flags, e.g. -DPFB_PROFILE -DPFB_CHASE=21]
set -e
NAME=$1; shift
WHAT=${1:-lgss}; shift || true
ROOT=$(cd "$(dirname "$0")/.." && pwd)
SRC=$ROOT/qnmh_sysid2018_b200/csrc
OUT=$ROOT/build_variants/$NAME
mkdir -p $OUT
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC $*"
if [ "$WHAT" = lgss ]; then UNITS="pfb200 pf_instances_lgss"; FLAGS="$FLAGS -DPFB_LGSS_ONLY"; else UNITS="pfb200 pf_instances_lgss pf_instances_sv pf_instances_svl pf_instances_lgss_fa"; fi
pids=""
for u in $UNITS; do
  ( cd $SRC && nvcc $FLAGS -c -o $OUT/$u.o $u.cu 2> $OUT/$u.log ) &
  pids="$pids $!"
done
for p in $pids; do wait $p; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $OUT/libpfb200.so $(for u in $UNITS; do echo $OUT/$u.o; done)
grep -h "error" $OUT/*.log || true
ls -la $OUT/libpfb200.so
