#!/bin/bash
# round 2, re-entry: whole GPU test tier, smoke, default bench, the workload lines, then the ncu
# launch list + one full capture of the C2 fp64 kernel (B200_PROFILING.md recipe)
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader
OUT=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee $OUT/r2_gputests.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 600 python bench.py > $OUT/r2_default_bench.json 2> $OUT/r2_default_bench.err; tail -c 2500 $OUT/r2_default_bench.json
for w in c1 c3 c4; do
  timeout 300 python bench.py --workload $w --no-cpu --steps 5 > $OUT/r2_${w}_f64_bench.json 2>> $OUT/r2_wl.err; tail -c 1200 $OUT/r2_${w}_f64_bench.json
done
timeout 300 python bench.py --dtype f32 --no-cpu --steps 5 > $OUT/r2_f32_bench.json 2>> $OUT/r2_wl.err; tail -c 1200 $OUT/r2_f32_bench.json
timeout 900 bash tools/profile_bench.sh r2 f64
