#!/bin/bash
# round-2 GPU session G: where do the fused BatchNorm-backward statistics pay?  (threshold on the conv's reduction depth)
set -x
O=gpurun_out
for K in 0 1500 2500 4000 1073741824; do
  SG_BSTATS_MIN_K=$K python bench.py --no-cpu-baseline --steps 10 > $O/g_bench_$K.log 2> $O/g_bench_$K.err
done
