echo "== base: no smoother"; timeout 120 python tools/run_once.py --T 200 --no-smoother 2>&1 | tail -1
for V in c21 c28; do
echo "== $V"
export PFB200_LIB=/root/repo/build_variants/$V/libpfb200.so
timeout 120 python tools/run_once.py --T 200 2>&1 | tail -1
timeout 120 python tools/run_once.py --T 200 --dtype f32 2>&1 | tail -1
timeout 120 python tools/run_once.py --T 100 --N 4096 --B 1024 2>&1 | tail -1
timeout 120 python tools/run_once.py --T 50 --N 16777216 2>&1 | tail -1
done
