timeout 300 python tools/sv_scan.py 2>&1 | tail -16
export PFB200_LIB=$PWD/build_variants/prof_all/libpfb200.so
[ -f $PFB200_LIB ] && timeout 300 python tools/sv_scan.py 2>&1 | tail -4
