for V in base; do
echo "== $V"
if [ $V != base ]; then export PFB200_LIB=/root/repo/build_variants/$V/libpfb200.so; fi
timeout 120 python tools/run_once.py --T 200 2>&1 | tail -1
timeout 120 python tools/run_once.py --T 200 --dtype f32 2>&1 | tail -1
timeout 120 python tools/run_once.py --T 100 --N 4096 --B 1024 2>&1 | tail -1
timeout 120 python tools/run_once.py --T 50 --N 16777216 2>&1 | tail -1
timeout 120 python tools/run_once.py --T 200 --N 65536 2>&1 | tail -1
done
