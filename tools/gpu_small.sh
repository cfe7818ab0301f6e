for V in base s4; do
echo "== $V"
if [ $V != base ]; then export PFB200_LIB=/root/repo/build_variants/$V/libpfb200.so; fi
for N in 300 600 1000 1500 2000; do
timeout 120 python tools/run_once.py --T 200 --N $N 2>&1 | tail -1
done
timeout 120 python tools/run_once.py --T 200 --N 1000 --B 512 2>&1 | tail -1
done
