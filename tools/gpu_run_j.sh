#!/bin/bash
# round-2 GPU session J: main chain captured on a high-priority stream?
set -x
O=gpurun_out
for P in 0 1 0 1; do
  SG_MAIN_PRIO=$P python bench.py --no-cpu-baseline --steps 20 > $O/j_bench_${P}_$RANDOM.log 2>> $O/j.err
done
grep -h -o '"ms_per_step": [0-9.]*' $O/j_bench_*.log
